/* `_jaxsplat` -- the compiled Python module the reference's host layer imports
 * (`import _jaxsplat`: jaxsplat/_project/impl.py, jaxsplat/_rasterize/impl.py,
 * jaxsplat/_project/lowering.py, jaxsplat/_rasterize/lowering.py), backed by libjaxsplat_b200.so.
 *
 * Replaces the pybind11 module of reference lib/ffi.cu:54-68 with the same two functions:
 *   registrations() -> {"project_fwd" | "project_bwd" | "rasterize_fwd" | "rasterize_bwd":
 *                       PyCapsule named "xla._CUSTOM_CALL_TARGET"}            (lib/ffi.cu:4-13)
 *   make_descriptor(num_points, img_shape, f, c, glob_scale, clip_thresh, block_width) -> bytes
 *                                                                             (lib/ffi.cu:15-52)
 * Plain CPython C API (no pybind11 needed to build it).  Extensions, keyword-only in spirit:
 *   registrations(status_returning=True) hands out the five-argument targets that report failures
 *   through XlaCustomCallStatus (include/jaxsplat_b200.h); last_error() / clear_error() expose the
 *   library's sticky error text.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>

#include "../../include/jaxsplat_b200.h"

static const char *kCapsuleName = "xla._CUSTOM_CALL_TARGET";

static int put_capsule(PyObject *dict, const char *name, void *fn) {
    PyObject *cap = PyCapsule_New(fn, kCapsuleName, NULL);
    if (!cap) return -1;
    const int rc = PyDict_SetItemString(dict, name, cap);
    Py_DECREF(cap);
    return rc;
}

static PyObject *py_registrations(PyObject *self, PyObject *args, PyObject *kwargs) {
    (void)self;
    static char *kwlist[] = {"status_returning", NULL};
    int status_returning = 0;
    if (!PyArg_ParseTupleAndKeywords(args, kwargs, "|p", kwlist, &status_returning)) return NULL;
    PyObject *dict = PyDict_New();
    if (!dict) return NULL;
    int rc;
    if (status_returning) {
        rc = put_capsule(dict, "project_fwd", (void *)&jsb200_project_fwd_status) |
             put_capsule(dict, "project_bwd", (void *)&jsb200_project_bwd_status) |
             put_capsule(dict, "rasterize_fwd", (void *)&jsb200_rasterize_fwd_status) |
             put_capsule(dict, "rasterize_bwd", (void *)&jsb200_rasterize_bwd_status);
    } else {
        rc = put_capsule(dict, "project_fwd", (void *)&jsb200_project_fwd) |
             put_capsule(dict, "project_bwd", (void *)&jsb200_project_bwd) |
             put_capsule(dict, "rasterize_fwd", (void *)&jsb200_rasterize_fwd) |
             put_capsule(dict, "rasterize_bwd", (void *)&jsb200_rasterize_bwd);
    }
    if (rc != 0) {
        Py_DECREF(dict);
        return NULL;
    }
    return dict;
}

static PyObject *py_make_descriptor(PyObject *self, PyObject *args, PyObject *kwargs) {
    (void)self;
    static char *kwlist[] = {"num_points", "img_shape", "f", "c", "glob_scale", "clip_thresh",
                             "block_width", NULL};
    unsigned int num_points = 0, img_h = 0, img_w = 0, block_width = 0;
    float fx = 0.f, fy = 0.f, cx = 0.f, cy = 0.f, glob_scale = 0.f, clip_thresh = 0.f;
    if (!PyArg_ParseTupleAndKeywords(args, kwargs, "I(II)(ff)(ff)ffI", kwlist, &num_points, &img_h,
                                     &img_w, &fx, &fy, &cx, &cy, &glob_scale, &clip_thresh,
                                     &block_width))
        return NULL;
    char out[JSB200_DESCRIPTOR_BYTES];
    if (jsb200_make_descriptor(num_points, img_h, img_w, fx, fy, cx, cy, glob_scale, clip_thresh,
                               block_width, out) != 0) {
        const char *msg = jsb200_last_error();
        PyErr_SetString(PyExc_ValueError, msg ? msg : "make_descriptor failed");
        jsb200_clear_error();
        return NULL;
    }
    return PyBytes_FromStringAndSize(out, JSB200_DESCRIPTOR_BYTES);
}

static PyObject *py_last_error(PyObject *self, PyObject *noargs) {
    (void)self;
    (void)noargs;
    const char *msg = jsb200_last_error();
    if (!msg) Py_RETURN_NONE;
    return PyUnicode_FromString(msg);
}

static PyObject *py_clear_error(PyObject *self, PyObject *noargs) {
    (void)self;
    (void)noargs;
    jsb200_clear_error();
    Py_RETURN_NONE;
}

static PyMethodDef kMethods[] = {
    {"registrations", (PyCFunction)(void (*)(void))py_registrations, METH_VARARGS | METH_KEYWORDS,
     "registrations(status_returning=False) -> dict of XLA custom-call target capsules"},
    {"make_descriptor", (PyCFunction)(void (*)(void))py_make_descriptor,
     METH_VARARGS | METH_KEYWORDS,
     "make_descriptor(num_points, img_shape, f, c, glob_scale, clip_thresh, block_width) -> bytes"},
    {"last_error", py_last_error, METH_NOARGS, "sticky error text of this thread's last failing call, or None"},
    {"clear_error", py_clear_error, METH_NOARGS, "reset the sticky error"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef kModule = {PyModuleDef_HEAD_INIT, "_jaxsplat",
                                     "jaxsplat native module backed by libjaxsplat_b200.so (sm_100a)",
                                     -1, kMethods, NULL, NULL, NULL, NULL};

PyMODINIT_FUNC PyInit__jaxsplat(void) { return PyModule_Create(&kModule); }
