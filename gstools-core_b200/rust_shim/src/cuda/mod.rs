//! FFI to libgsfield (include/gsfield.h).  UNBUILT in this repository; see ../../README.md.
//!
//! ndarray strides are already in elements, which is what the C ABI takes, so arbitrary views
//! (`PyReadonlyArray::as_array()`, src/lib.rs:43-46) are passed without copying.
use std::ffi::CStr;
use std::os::raw::{c_char, c_int};

use ndarray::{Array1, Array2, ArrayView1, ArrayView2, ShapeBuilder};

extern "C" {
    fn gsf_summate(
        dim: c_int, n_modes: i64, n_points: i64,
        cov: *const f64, cov_s0: i64, cov_s1: i64,
        z1: *const f64, z1_s: i64, z2: *const f64, z2_s: i64,
        pos: *const f64, pos_s0: i64, pos_s1: i64,
        out: *mut f64, num_threads: c_int,
    ) -> c_int;
    fn gsf_summate_incompr(
        dim: c_int, n_modes: i64, n_points: i64,
        cov: *const f64, cov_s0: i64, cov_s1: i64,
        z1: *const f64, z1_s: i64, z2: *const f64, z2_s: i64,
        pos: *const f64, pos_s0: i64, pos_s1: i64,
        out: *mut f64, out_s0: i64, out_s1: i64, num_threads: c_int,
    ) -> c_int;
    fn gsf_summate_fourier(
        dim: c_int, n_modes: i64, n_points: i64,
        sf: *const f64, sf_s: i64,
        modes: *const f64, modes_s0: i64, modes_s1: i64,
        z1: *const f64, z1_s: i64, z2: *const f64, z2_s: i64,
        pos: *const f64, pos_s0: i64, pos_s1: i64,
        out: *mut f64, num_threads: c_int,
    ) -> c_int;
    fn gsf_last_error() -> *const c_char;
}

fn check(rc: c_int) {
    if rc != 0 {
        let msg = unsafe { CStr::from_ptr(gsf_last_error()) }.to_string_lossy().into_owned();
        // the reference's error convention is panic (asserts, src/field.rs:44-46,180)
        panic!("gsfield: {msg} [status {rc}]");
    }
}

fn s2(a: &ArrayView2<'_, f64>) -> (i64, i64) {
    (a.strides()[0] as i64, a.strides()[1] as i64)
}

pub fn summate(
    cov: ArrayView2<'_, f64>, z1: ArrayView1<'_, f64>, z2: ArrayView1<'_, f64>,
    pos: ArrayView2<'_, f64>, num_threads: Option<usize>,
) -> Array1<f64> {
    let (d, n) = cov.dim();
    let m = pos.dim().1;
    let mut out = Array1::<f64>::zeros(m);
    let (c0, c1) = s2(&cov);
    let (p0, p1) = s2(&pos);
    check(unsafe {
        gsf_summate(
            d as c_int, n as i64, m as i64, cov.as_ptr(), c0, c1,
            z1.as_ptr(), z1.strides()[0] as i64, z2.as_ptr(), z2.strides()[0] as i64,
            pos.as_ptr(), p0, p1, out.as_mut_ptr(), num_threads.unwrap_or(0) as c_int,
        )
    });
    out
}

pub fn summate_incompr(
    cov: ArrayView2<'_, f64>, z1: ArrayView1<'_, f64>, z2: ArrayView1<'_, f64>,
    pos: ArrayView2<'_, f64>, num_threads: Option<usize>,
) -> Array2<f64> {
    let (d, n) = cov.dim();
    let m = pos.dim().1;
    // (d, M) in Fortran order == the reference's from_shape_vec((M, N)).reversed_axes(), :166-174
    let mut out = Array2::<f64>::zeros((d, m).f());
    let (c0, c1) = s2(&cov);
    let (p0, p1) = s2(&pos);
    let (o0, o1) = (out.strides()[0] as i64, out.strides()[1] as i64);
    check(unsafe {
        gsf_summate_incompr(
            d as c_int, n as i64, m as i64, cov.as_ptr(), c0, c1,
            z1.as_ptr(), z1.strides()[0] as i64, z2.as_ptr(), z2.strides()[0] as i64,
            pos.as_ptr(), p0, p1, out.as_mut_ptr(), o0, o1, num_threads.unwrap_or(0) as c_int,
        )
    });
    out
}

pub fn summate_fourier(
    sf: ArrayView1<'_, f64>, modes: ArrayView2<'_, f64>, z1: ArrayView1<'_, f64>,
    z2: ArrayView1<'_, f64>, pos: ArrayView2<'_, f64>, num_threads: Option<usize>,
) -> Array1<f64> {
    assert_eq!(sf.dim(), modes.dim().1); // the reference fails inside Zip::and instead
    let (d, n) = modes.dim();
    let m = pos.dim().1;
    let mut out = Array1::<f64>::zeros(m);
    let (c0, c1) = s2(&modes);
    let (p0, p1) = s2(&pos);
    check(unsafe {
        gsf_summate_fourier(
            d as c_int, n as i64, m as i64, sf.as_ptr(), sf.strides()[0] as i64,
            modes.as_ptr(), c0, c1,
            z1.as_ptr(), z1.strides()[0] as i64, z2.as_ptr(), z2.strides()[0] as i64,
            pos.as_ptr(), p0, p1, out.as_mut_ptr(), num_threads.unwrap_or(0) as c_int,
        )
    });
    out
}
