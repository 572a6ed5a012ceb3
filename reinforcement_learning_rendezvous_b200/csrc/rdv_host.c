/* rdv_host.c -- CPython helper of RendezvousVecEnv (host side of the drop-in boundary, no device code).
 *
 * SB3's VecEnv contract (DummyVecEnv.step_wait + Monitor.step, /root/reference/main.py:33-34) wants one Python dict per
 * env and step, and for every env whose episode ended {"terminal_observation": ndarray, "episode": {"r", "l", "t"}}.
 * At 65,536 envs and ~3,300 episode ends per step the interpreter spends more time creating those objects than the
 * GPU spends stepping the envs; this module builds them from the finished rows (RdvFinishedRow, include/rdv_b200.h)
 * with direct C-API calls.  Built by _native.build() with gcc against Python.h and numpy's headers (the row views of
 * the [m,17] terminal-observation array are made with PyArray_NewFromDescr).
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#define NPY_NO_DEPRECATED_API NPY_1_7_API_VERSION
#include <numpy/arrayobject.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#define PF_FAR 16          /* iterations ahead: the slot's dict object ... */
#define PF_NEAR 8          /* ... and its key table */

typedef struct {
    int32_t env, end_reason;
    float terminal_obs[17];
    float pad;
    double record[6];           /* RDV_EP_RETURN, LENGTH, SUCCESS, COLLIDED, DELTA_V, DELTA_W */
} FinishedRow;

static PyObject *k_term, *k_episode, *k_r, *k_l, *k_t, *k_success, *k_collided, *k_dv, *k_dw, *k_reason;

/* build_infos(infos: list, dirty: buffer of int32, rows: buffer of m 128-byte rows, term: [m,17] array, elapsed: float,
 *             rich: bool, reasons: tuple[str]) -> bytes (int32[m])
 * Slots named by `dirty` (last step's episode-end dicts) become empty dicts; slot rows[j].env gets the
 * episode-end dict of row j.  A slot's dict is reused when the list holds the only reference to it and replaced by a
 * new one otherwise, so a dict the caller kept is never changed behind its back.  Returns the env indices of this
 * step's rows as packed int32 (the next call's `dirty`; one object instead of a list of m Python ints). */
static PyObject *build_infos(PyObject *self, PyObject *args)
{
    PyObject *infos, *dirty_obj, *rows_obj, *term, *reasons;
    double elapsed;
    int rich;
    if (!PyArg_ParseTuple(args, "O!OOOdpO!", &PyList_Type, &infos, &dirty_obj, &rows_obj, &term, &elapsed,
                          &rich, &PyTuple_Type, &reasons))
        return NULL;
    Py_buffer dview;
    if (PyObject_GetBuffer(dirty_obj, &dview, PyBUF_SIMPLE) < 0) return NULL;
    if (dview.len % (Py_ssize_t)sizeof(int32_t)) {
        PyBuffer_Release(&dview);
        PyErr_SetString(PyExc_ValueError, "dirty must hold packed int32 env indices");
        return NULL;
    }
    const int32_t *dirty = (const int32_t *)dview.buf;
    const Py_ssize_t nd = dview.len / (Py_ssize_t)sizeof(int32_t);
    const Py_ssize_t n = PyList_GET_SIZE(infos);
    /* The 65,536 per-env dicts are spread over ~20 MB of heap and the finished envs are a random subset of them: left
     * alone, every slot costs two or three cache misses (the dict, its key table).  The slots a few iterations ahead
     * are prefetched -- first the dict object, then, once that has arrived, its key table. */
    for (Py_ssize_t k = 0; k < nd; ++k) {
        if (k + PF_FAR < nd) {
            const Py_ssize_t f = dirty[k + PF_FAR];
            if (f >= 0 && f < n) __builtin_prefetch(PyList_GET_ITEM(infos, f));
        }
        if (k + PF_NEAR < nd) {
            const Py_ssize_t f = dirty[k + PF_NEAR];
            if (f >= 0 && f < n) {
                PyObject *o = PyList_GET_ITEM(infos, f);
                if (PyDict_CheckExact(o)) {
                    const char *kp = (const char *)((PyDictObject *)o)->ma_keys;
                    __builtin_prefetch(kp); __builtin_prefetch(kp + 64); __builtin_prefetch(kp + 128);
                }
            }
        }
        const Py_ssize_t i = dirty[k];
        if (i < 0 || i >= n) {
            PyBuffer_Release(&dview);
            PyErr_SetString(PyExc_IndexError, "dirty index out of range");
            return NULL;
        }
        /* a dict nobody else holds is emptied in place (no one can tell it from a fresh one); one the caller kept a
         * reference to is left alone and replaced */
        PyObject *old = PyList_GET_ITEM(infos, i);
        if (PyDict_CheckExact(old) && Py_REFCNT(old) == 1) {
            PyDict_Clear(old);
        } else {
            PyObject *d = PyDict_New();
            if (!d) { PyBuffer_Release(&dview); return NULL; }
            PyList_SetItem(infos, i, d);                             /* steals d, releases the old dict */
        }
    }
    PyBuffer_Release(&dview);
    Py_buffer view;
    if (PyObject_GetBuffer(rows_obj, &view, PyBUF_SIMPLE) < 0) return NULL;
    if (view.len % (Py_ssize_t)sizeof(FinishedRow)) {
        PyBuffer_Release(&view);
        PyErr_SetString(PyExc_ValueError, "rows must hold whole 128-byte RdvFinishedRow records");
        return NULL;
    }
    const Py_ssize_t m = view.len / (Py_ssize_t)sizeof(FinishedRow);
    const FinishedRow *rows = (const FinishedRow *)view.buf;
    if (!PyArray_Check(term) || PyArray_TYPE((PyArrayObject *)term) != NPY_FLOAT32 || PyArray_NDIM((PyArrayObject *)term) != 2 ||
        PyArray_DIM((PyArrayObject *)term, 1) != 17 || PyArray_DIM((PyArrayObject *)term, 0) < m ||
        !PyArray_IS_C_CONTIGUOUS((PyArrayObject *)term)) {
        PyBuffer_Release(&view);
        PyErr_SetString(PyExc_TypeError, "term must be a C-contiguous float32 [m,17] array");
        return NULL;
    }
    char *term_data = PyArray_BYTES((PyArrayObject *)term);
    PyArray_Descr *f32_descr = PyArray_DESCR((PyArrayObject *)term);
    PyObject *out = PyBytes_FromStringAndSize(NULL, m * (Py_ssize_t)sizeof(int32_t));
    PyObject *t_obj = PyFloat_FromDouble(elapsed);
    /* this step's episode dicts are clones of one template {"r", "l", "t": elapsed}: a clone copies the key table in
     * one piece, and writing "r" / "l" replaces values in place (no insertion, no resize) */
    PyObject *ep_tmpl = PyDict_New();
    if (!out || !t_obj || !ep_tmpl) goto fail;
    int32_t *out_idx = (int32_t *)PyBytes_AS_STRING(out);
    if (PyDict_SetItem(ep_tmpl, k_r, Py_None) < 0 || PyDict_SetItem(ep_tmpl, k_l, Py_None) < 0 ||
        PyDict_SetItem(ep_tmpl, k_t, t_obj) < 0)
        goto fail;
    for (Py_ssize_t j = 0; j < m; ++j) {
        if (j + PF_FAR < m) {
            const int32_t f = rows[j + PF_FAR].env;
            if (f >= 0 && f < n) __builtin_prefetch(PyList_GET_ITEM(infos, f));
        }
        FinishedRow row;
        memcpy(&row, rows + j, sizeof(row));
        if (row.env < 0 || row.env >= n) {
            PyErr_SetString(PyExc_IndexError, "finished row names an env outside the batch");
            goto fail;
        }
        /* row view of the caller's own [m,17] array (what term[j] returns, without the generic indexing path) */
        npy_intp dim = 17;
        Py_INCREF(f32_descr);
        PyObject *obs = PyArray_NewFromDescr(&PyArray_Type, f32_descr, 1, &dim, NULL,
                                             term_data + (size_t)j * 17 * sizeof(float), NPY_ARRAY_CARRAY, NULL);
        if (obs) {
            Py_INCREF(term);
            if (PyArray_SetBaseObject((PyArrayObject *)obs, term) < 0) { Py_DECREF(obs); obs = NULL; }   /* steals term */
        }
        PyObject *slot = PyList_GET_ITEM(infos, row.env);
        const int reuse = PyDict_CheckExact(slot) && Py_REFCNT(slot) == 1;   /* the env's own dict, held by nobody else */
        PyObject *ep = PyDict_Copy(ep_tmpl), *d = reuse ? slot : PyDict_New();
        if (reuse && PyDict_GET_SIZE(d)) PyDict_Clear(d);
        PyObject *r = PyFloat_FromDouble(rint(row.record[0] * 1e6) / 1e6);     /* Monitor rounds the return to 6 places */
        PyObject *l = PyLong_FromLong((long)row.record[1]);
        int bad = !obs || !ep || !d || !r || !l;
        if (!bad) {
            bad |= PyDict_SetItem(ep, k_r, r) < 0 || PyDict_SetItem(ep, k_l, l) < 0;
            bad |= PyDict_SetItem(d, k_term, obs) < 0 || PyDict_SetItem(d, k_episode, ep) < 0;
        }
        if (!bad && rich) {
            PyObject *dv = PyFloat_FromDouble(row.record[4]), *dw = PyFloat_FromDouble(row.record[5]);
            PyObject *why = (row.end_reason >= 0 && row.end_reason < PyTuple_GET_SIZE(reasons))
                                ? PyTuple_GET_ITEM(reasons, row.end_reason) : Py_None;
            bad |= !dv || !dw;
            if (!bad) {
                bad |= PyDict_SetItem(d, k_success, row.record[2] > 0 ? Py_True : Py_False) < 0;
                bad |= PyDict_SetItem(d, k_collided, row.record[3] > 0 ? Py_True : Py_False) < 0;
                bad |= PyDict_SetItem(d, k_dv, dv) < 0 || PyDict_SetItem(d, k_dw, dw) < 0;
                bad |= PyDict_SetItem(d, k_reason, why) < 0;
            }
            Py_XDECREF(dv); Py_XDECREF(dw);
        }
        Py_XDECREF(obs); Py_XDECREF(ep); Py_XDECREF(r); Py_XDECREF(l);
        if (bad) { if (!reuse) Py_XDECREF(d); goto fail; }
        if (!reuse) PyList_SetItem(infos, row.env, d);               /* steals d */
        out_idx[j] = row.env;
    }
    Py_DECREF(t_obj);
    Py_DECREF(ep_tmpl);
    PyBuffer_Release(&view);
    return out;
fail:
    Py_XDECREF(t_obj);
    Py_XDECREF(ep_tmpl);
    Py_XDECREF(out);
    PyBuffer_Release(&view);
    return NULL;
}

static PyMethodDef methods[] = {
    {"build_infos", build_infos, METH_VARARGS, "Per-env info dicts of one VecEnv step from the finished rows."},
    {NULL, NULL, 0, NULL},
};
static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_rdv_host", "host-side helpers of RendezvousVecEnv", -1, methods};

PyMODINIT_FUNC PyInit__rdv_host(void)
{
    import_array();
    k_term = PyUnicode_InternFromString("terminal_observation");
    k_episode = PyUnicode_InternFromString("episode");
    k_r = PyUnicode_InternFromString("r");
    k_l = PyUnicode_InternFromString("l");
    k_t = PyUnicode_InternFromString("t");
    k_success = PyUnicode_InternFromString("is_success");
    k_collided = PyUnicode_InternFromString("collided");
    k_dv = PyUnicode_InternFromString("total_delta_v");
    k_dw = PyUnicode_InternFromString("total_delta_w");
    k_reason = PyUnicode_InternFromString("end_reason");
    return PyModule_Create(&module);
}
