/* rdv_host.c -- CPython helper of RendezvousVecEnv (host side of the drop-in boundary, no device code).
 *
 * SB3's VecEnv contract (DummyVecEnv.step_wait + Monitor.step, /root/reference/main.py:33-34) wants one Python dict per
 * env and step, and for every env whose episode ended {"terminal_observation": ndarray, "episode": {"r", "l", "t"}}.
 * At 65,536 envs and ~3,300 episode ends per step the interpreter spends more time creating those objects than the
 * GPU spends stepping the envs; this module builds them from the finished rows (RdvFinishedRow, include/rdv_b200.h)
 * with direct C-API calls.  Built by _native.build() with gcc against Python.h; no numpy C API is needed (row views
 * come from the sequence protocol of the [m,17] array the caller passes).
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct {
    int32_t env, end_reason;
    float terminal_obs[17];
    float pad;
    double record[6];           /* RDV_EP_RETURN, LENGTH, SUCCESS, COLLIDED, DELTA_V, DELTA_W */
} FinishedRow;

static PyObject *k_term, *k_episode, *k_r, *k_l, *k_t, *k_success, *k_collided, *k_dv, *k_dw, *k_reason;

/* build_infos(infos: list, dirty: list[int], rows: buffer of m 128-byte rows, term: [m,17] array, elapsed: float,
 *             rich: bool, reasons: tuple[str]) -> list[int]
 * Slots named by `dirty` (last step's episode-end dicts) become empty dicts; slot rows[j].env gets the
 * episode-end dict of row j.  A slot's dict is reused when the list holds the only reference to it and replaced by a
 * new one otherwise, so a dict the caller kept is never changed behind its back.  Returns the env indices of this step's rows (the next call's `dirty`). */
static PyObject *build_infos(PyObject *self, PyObject *args)
{
    PyObject *infos, *dirty, *rows_obj, *term, *reasons;
    double elapsed;
    int rich;
    if (!PyArg_ParseTuple(args, "O!O!OOdpO!", &PyList_Type, &infos, &PyList_Type, &dirty, &rows_obj, &term, &elapsed,
                          &rich, &PyTuple_Type, &reasons))
        return NULL;
    const Py_ssize_t n = PyList_GET_SIZE(infos);
    for (Py_ssize_t k = 0; k < PyList_GET_SIZE(dirty); ++k) {
        const Py_ssize_t i = PyLong_AsSsize_t(PyList_GET_ITEM(dirty, k));
        if (i < 0 || i >= n) {
            if (!PyErr_Occurred()) PyErr_SetString(PyExc_IndexError, "dirty index out of range");
            return NULL;
        }
        /* a dict nobody else holds is emptied in place (no one can tell it from a fresh one); one the caller kept a
         * reference to is left alone and replaced */
        PyObject *old = PyList_GET_ITEM(infos, i);
        if (PyDict_CheckExact(old) && Py_REFCNT(old) == 1) {
            PyDict_Clear(old);
        } else {
            PyObject *d = PyDict_New();
            if (!d) return NULL;
            PyList_SetItem(infos, i, d);                             /* steals d, releases the old dict */
        }
    }
    Py_buffer view;
    if (PyObject_GetBuffer(rows_obj, &view, PyBUF_SIMPLE) < 0) return NULL;
    if (view.len % (Py_ssize_t)sizeof(FinishedRow)) {
        PyBuffer_Release(&view);
        PyErr_SetString(PyExc_ValueError, "rows must hold whole 128-byte RdvFinishedRow records");
        return NULL;
    }
    const Py_ssize_t m = view.len / (Py_ssize_t)sizeof(FinishedRow);
    const FinishedRow *rows = (const FinishedRow *)view.buf;
    PyObject *out = PyList_New(m);
    PyObject *t_obj = PyFloat_FromDouble(elapsed);
    if (!out || !t_obj) goto fail;
    for (Py_ssize_t j = 0; j < m; ++j) {
        FinishedRow row;
        memcpy(&row, rows + j, sizeof(row));
        if (row.env < 0 || row.env >= n) {
            PyErr_SetString(PyExc_IndexError, "finished row names an env outside the batch");
            goto fail;
        }
        PyObject *obs = PySequence_GetItem(term, j);                 /* row view of the caller's own [m,17] array */
        PyObject *slot = PyList_GET_ITEM(infos, row.env);
        const int reuse = PyDict_CheckExact(slot) && Py_REFCNT(slot) == 1;   /* the env's own dict, held by nobody else */
        PyObject *ep = PyDict_New(), *d = reuse ? slot : PyDict_New();
        if (reuse && PyDict_GET_SIZE(d)) PyDict_Clear(d);
        PyObject *r = PyFloat_FromDouble(rint(row.record[0] * 1e6) / 1e6);     /* Monitor rounds the return to 6 places */
        PyObject *l = PyLong_FromLong((long)row.record[1]);
        int bad = !obs || !ep || !d || !r || !l;
        if (!bad) {
            bad |= PyDict_SetItem(ep, k_r, r) < 0 || PyDict_SetItem(ep, k_l, l) < 0 || PyDict_SetItem(ep, k_t, t_obj) < 0;
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
        PyObject *idx = PyLong_FromLong(row.env);
        if (!idx) goto fail;
        PyList_SET_ITEM(out, j, idx);
    }
    Py_DECREF(t_obj);
    PyBuffer_Release(&view);
    return out;
fail:
    Py_XDECREF(t_obj);
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
