/* gsf_pybind.c -- native CPython binding of the three summation entry points.
 *
 * The reference's Python surface is a native pyo3 module (/root/reference/src/lib.rs:33-84): argument
 * conversion costs well under a microsecond there.  The ctypes mirror in gstools_core/__init__.py
 * spends ~20 us per call on marshalling alone, which is most of a C1-sized call (100 modes x 1e4
 * points, ~25 us of GPU + runtime).  This module is the fast lane for exactly that case:
 *
 *     summate(cov, z1, z2, pos, num_threads)            -> ndarray | int status | None
 *     summate_incompr(cov, z1, z2, pos, num_threads)    -> ndarray | int status | None
 *     summate_fourier(sf, modes, z1, z2, pos, num_threads) -> ndarray | int status | None
 *
 * None means "not handled here" (an argument is not a float64 buffer, shapes disagree, the result
 * is large enough for the pinned-pool path, ...): the Python wrapper then runs its general path,
 * which also produces the reference-shaped exceptions.  An int is a non-zero gsf_status for the
 * wrapper to raise.  No compute happens here -- the work is gsf_summate* in libgsfield.so, whose
 * addresses the wrapper hands over with bind() (so GSF_LIB overrides stay coherent).
 *
 * Only the buffer protocol is used (no numpy headers): numpy arrays export format "d" with shape
 * and byte strides; the result array is made by calling numpy.empty.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <string.h>

typedef int (*summate_fn)(int, int64_t, int64_t, const double *, int64_t, int64_t, const double *, int64_t,
                          const double *, int64_t, const double *, int64_t, int64_t, double *, int);
typedef int (*incompr_fn)(int, int64_t, int64_t, const double *, int64_t, int64_t, const double *, int64_t,
                          const double *, int64_t, const double *, int64_t, int64_t, double *, int64_t, int64_t, int);
typedef int (*fourier_fn)(int, int64_t, int64_t, const double *, int64_t, const double *, int64_t, int64_t,
                          const double *, int64_t, const double *, int64_t, const double *, int64_t, int64_t,
                          double *, int);

static summate_fn g_summate;
static incompr_fn g_incompr;
static fourier_fn g_fourier;
static PyObject *g_ndarray;      /* numpy.ndarray: only real arrays take the fast lane */
static PyObject *g_empty;        /* numpy.empty */
static PyObject *g_kw_c, *g_kw_f; /* {"dtype": float64} / {"dtype": float64, "order": "F"} */
static Py_ssize_t g_max_bytes = 256 * 1024;   /* larger results take the wrapper's pinned-pool path */

typedef struct {
    Py_buffer v;
    int held;
} Arg;

/* float64 buffer of `ndim` dimensions with 8-byte-multiple strides; 0 if not (no exception set) */
static int take(PyObject *o, int ndim, Arg *a)
{
    a->held = 0;
    if (!g_ndarray || !PyObject_TypeCheck(o, (PyTypeObject *)g_ndarray)) return 0;
    if (PyObject_GetBuffer(o, &a->v, PyBUF_STRIDES | PyBUF_FORMAT) != 0) {
        PyErr_Clear();
        return 0;
    }
    a->held = 1;
    const char *f = a->v.format;
    if (!f || a->v.itemsize != 8 || a->v.ndim != ndim) return 0;
    if (f[0] == '<' || f[0] == '=' || f[0] == '@') ++f;
    if (f[0] != 'd' || f[1] != 0) return 0;
    for (int i = 0; i < ndim; ++i)
        if (a->v.strides[i] % 8) return 0;
    return 1;
}

static void drop(Arg *a, int n)
{
    for (int i = 0; i < n; ++i)
        if (a[i].held) PyBuffer_Release(&a[i].v);
}

/* num_threads: None -> 0, int >= 0; negative -> OverflowError like Option<usize> (src/lib.rs:41) */
static int threads_of(PyObject *o, int *out)
{
    *out = 0;
    if (o == NULL || o == Py_None) return 1;
    long v = PyLong_AsLong(o);
    if (v == -1 && PyErr_Occurred()) return 0;
    if (v < 0) {
        PyErr_SetString(PyExc_OverflowError, "can't convert negative int to unsigned");
        return 0;
    }
    *out = v > 0x7fffffffL ? 0x7fffffff : (int)v;
    return 1;
}

static PyObject *make_result(Py_ssize_t d, Py_ssize_t m, Arg *out)
{
    PyObject *shape = d > 0 ? Py_BuildValue("((nn))", d, m) : Py_BuildValue("((n))", m);
    if (!shape) return NULL;
    PyObject *arr = PyObject_Call(g_empty, shape, d > 0 ? g_kw_f : g_kw_c);
    Py_DECREF(shape);
    if (!arr) return NULL;
    out->held = 0;
    if (PyObject_GetBuffer(arr, &out->v, PyBUF_STRIDES | PyBUF_WRITABLE) != 0) {
        Py_DECREF(arr);
        return NULL;
    }
    out->held = 1;
    return arr;
}

#define S0(a) ((int64_t)((a).v.strides[0] / 8))
#define S1(a) ((int64_t)((a).v.strides[1] / 8))

/* kind: 0 summate, 1 incompr, 2 fourier */
static PyObject *run(int kind, PyObject *const *args, Py_ssize_t nargs)
{
    const int first = kind == 2 ? 1 : 0;
    if (nargs < first + 4 || nargs > first + 5) {
        PyErr_SetString(PyExc_TypeError, "wrong number of arguments");
        return NULL;
    }
    if ((kind == 0 && !g_summate) || (kind == 1 && !g_incompr) || (kind == 2 && !g_fourier)) Py_RETURN_NONE;
    int nthreads;
    if (!threads_of(nargs > first + 4 ? args[first + 4] : NULL, &nthreads)) return NULL;
    Arg a[6];
    memset(a, 0, sizeof a);
    /* a[0] cov (d, N), a[1] z1, a[2] z2, a[3] pos (d, M), a[4] sf, a[5] out */
    int ok = take(args[first], 2, &a[0]) && take(args[first + 1], 1, &a[1]) && take(args[first + 2], 1, &a[2]) &&
             take(args[first + 3], 2, &a[3]) && (kind != 2 || take(args[0], 1, &a[4]));
    if (ok) {
        const Py_ssize_t d = a[0].v.shape[0], n = a[0].v.shape[1], m = a[3].v.shape[1];
        ok = a[3].v.shape[0] == d && a[1].v.shape[0] == n && a[2].v.shape[0] == n && d >= 1 && d <= 64 &&
             (kind != 2 || a[4].v.shape[0] == n) && m >= 1 && n >= 1 &&
             (kind == 1 ? d : 1) * m * 8 < g_max_bytes;
    }
    if (!ok) {
        drop(a, 6);
        Py_RETURN_NONE;
    }
    const Py_ssize_t d = a[0].v.shape[0], n = a[0].v.shape[1], m = a[3].v.shape[1];
    PyObject *res = make_result(kind == 1 ? d : 0, m, &a[5]);
    if (!res) {
        drop(a, 6);
        return NULL;
    }
    int rc;
    Py_BEGIN_ALLOW_THREADS
    if (kind == 0)
        rc = g_summate((int)d, n, m, a[0].v.buf, S0(a[0]), S1(a[0]), a[1].v.buf, S0(a[1]), a[2].v.buf, S0(a[2]),
                       a[3].v.buf, S0(a[3]), S1(a[3]), a[5].v.buf, nthreads);
    else if (kind == 1)
        rc = g_incompr((int)d, n, m, a[0].v.buf, S0(a[0]), S1(a[0]), a[1].v.buf, S0(a[1]), a[2].v.buf, S0(a[2]),
                       a[3].v.buf, S0(a[3]), S1(a[3]), a[5].v.buf, S0(a[5]), S1(a[5]), nthreads);
    else
        rc = g_fourier((int)d, n, m, a[4].v.buf, S0(a[4]), a[0].v.buf, S0(a[0]), S1(a[0]), a[1].v.buf, S0(a[1]),
                       a[2].v.buf, S0(a[2]), a[3].v.buf, S0(a[3]), S1(a[3]), a[5].v.buf, nthreads);
    Py_END_ALLOW_THREADS
    drop(a, 6);
    if (rc != 0) {
        Py_DECREF(res);
        return PyLong_FromLong(rc);
    }
    return res;
}

static PyObject *py_summate(PyObject *self, PyObject *const *args, Py_ssize_t nargs)
{
    (void)self;
    return run(0, args, nargs);
}
static PyObject *py_incompr(PyObject *self, PyObject *const *args, Py_ssize_t nargs)
{
    (void)self;
    return run(1, args, nargs);
}
static PyObject *py_fourier(PyObject *self, PyObject *const *args, Py_ssize_t nargs)
{
    (void)self;
    return run(2, args, nargs);
}

/* bind(addr_summate, addr_incompr, addr_fourier, numpy.ndarray, numpy.empty, numpy.float64, max_result_bytes) */
static PyObject *py_bind(PyObject *self, PyObject *args)
{
    (void)self;
    unsigned long long f0, f1, f2;
    PyObject *ndarray, *empty, *f64;
    Py_ssize_t max_bytes;
    if (!PyArg_ParseTuple(args, "KKKOOOn", &f0, &f1, &f2, &ndarray, &empty, &f64, &max_bytes)) return NULL;
    if (!PyType_Check(ndarray)) {
        PyErr_SetString(PyExc_TypeError, "bind: numpy.ndarray type expected");
        return NULL;
    }
    PyObject *kw_c = Py_BuildValue("{s:O}", "dtype", f64);
    PyObject *kw_f = Py_BuildValue("{s:O,s:s}", "dtype", f64, "order", "F");
    if (!kw_c || !kw_f) {
        Py_XDECREF(kw_c);
        Py_XDECREF(kw_f);
        return NULL;
    }
    Py_INCREF(empty);
    Py_INCREF(ndarray);
    Py_XDECREF(g_ndarray);
    g_ndarray = ndarray;
    Py_XDECREF(g_empty);
    Py_XDECREF(g_kw_c);
    Py_XDECREF(g_kw_f);
    g_empty = empty;
    g_kw_c = kw_c;
    g_kw_f = kw_f;
    g_max_bytes = max_bytes;
    g_summate = (summate_fn)(uintptr_t)f0;
    g_incompr = (incompr_fn)(uintptr_t)f1;
    g_fourier = (fourier_fn)(uintptr_t)f2;
    Py_RETURN_NONE;
}

static PyMethodDef methods[] = {
    {"summate", (PyCFunction)(void (*)(void))py_summate, METH_FASTCALL, "fast lane of gstools_core.summate"},
    {"summate_incompr", (PyCFunction)(void (*)(void))py_incompr, METH_FASTCALL, "fast lane of gstools_core.summate_incompr"},
    {"summate_fourier", (PyCFunction)(void (*)(void))py_fourier, METH_FASTCALL, "fast lane of gstools_core.summate_fourier"},
    {"bind", py_bind, METH_VARARGS, "bind(addr_summate, addr_incompr, addr_fourier, numpy.ndarray, numpy.empty, numpy.float64, max_bytes)"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moduledef = {PyModuleDef_HEAD_INIT, "_gsf_native", "native binding of libgsfield (small-call fast lane)",
                                       -1, methods, NULL, NULL, NULL, NULL};

PyMODINIT_FUNC PyInit__gsf_native(void) { return PyModule_Create(&moduledef); }
