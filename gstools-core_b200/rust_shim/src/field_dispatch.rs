//! New bodies for the three public functions of the reference's src/field.rs.  Signatures are
//! unchanged (src/field.rs:37-43, 97-103, 219-226); the shape asserts stay so the panic
//! messages match; the Rayon implementations move to `cpu_reference::*` (timed baseline only).
//! UNBUILT in this repository; see ../README.md.
use ndarray::{Array1, Array2, ArrayView1, ArrayView2};

use crate::cuda;

pub fn summator(
    cov_samples: ArrayView2<'_, f64>,
    z1: ArrayView1<'_, f64>,
    z2: ArrayView1<'_, f64>,
    pos: ArrayView2<'_, f64>,
    num_threads: Option<usize>,
) -> Array1<f64> {
    assert_eq!(cov_samples.dim().0, pos.dim().0);
    assert_eq!(cov_samples.dim().1, z1.dim());
    assert_eq!(cov_samples.dim().1, z2.dim());
    cuda::summate(cov_samples, z1, z2, pos, num_threads)
}

pub fn summator_incompr(
    cov_samples: ArrayView2<'_, f64>,
    z1: ArrayView1<'_, f64>,
    z2: ArrayView1<'_, f64>,
    pos: ArrayView2<'_, f64>,
    num_threads: Option<usize>,
) -> Array2<f64> {
    assert_eq!(cov_samples.dim().0, pos.dim().0);
    assert_eq!(cov_samples.dim().1, z1.dim());
    assert_eq!(cov_samples.dim().1, z2.dim());
    match pos.dim().0 {
        2 | 3 => cuda::summate_incompr(cov_samples, z1, z2, pos, num_threads),
        _ => panic!("Only two- and three-dimensional problems are supported."),
    }
}

pub fn summator_fourier(
    spectrum_factor: ArrayView1<'_, f64>,
    modes: ArrayView2<'_, f64>,
    z1: ArrayView1<'_, f64>,
    z2: ArrayView1<'_, f64>,
    pos: ArrayView2<'_, f64>,
    num_threads: Option<usize>,
) -> Array1<f64> {
    assert_eq!(modes.dim().0, pos.dim().0);
    assert_eq!(modes.dim().1, z1.dim());
    assert_eq!(modes.dim().1, z2.dim());
    cuda::summate_fourier(spectrum_factor, modes, z1, z2, pos, num_threads)
}
