/* packer.c — CPython extension: walk a list of molgym ObservationType tuples once and write flat arrays.
 *
 * Replaces the per-observation Python work of the reference's parse path (molgym/spaces.py:55-61,106-107;
 * molgym/agents/covariant/tools.py:8-49; agent.py:165-197): observation = (canvas, bag),
 * canvas = tuple of canvas_size items (label_index, (x, y, z)), bag = tuple of counts aligned with zs.
 *
 *   flatten(observations, canvas_size, num_species, labels:int32[B,N], xyz:float64[B,N,3], bags:float32[B,Z]) -> None
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

static int get_buf(PyObject* o, Py_buffer* b, Py_ssize_t need_bytes, const char* what) {
  if (PyObject_GetBuffer(o, b, PyBUF_WRITABLE | PyBUF_C_CONTIGUOUS) != 0) return -1;
  if (b->len < need_bytes) {
    PyErr_Format(PyExc_ValueError, "%s buffer too small: %zd < %zd bytes", what, b->len, need_bytes);
    PyBuffer_Release(b);
    return -1;
  }
  return 0;
}

static PyObject* flatten(PyObject* self, PyObject* args) {
  PyObject *observations, *o_labels, *o_xyz, *o_bags;
  Py_ssize_t N, Z;
  if (!PyArg_ParseTuple(args, "OnnOOO", &observations, &N, &Z, &o_labels, &o_xyz, &o_bags)) return NULL;
  PyObject* seq = PySequence_Fast(observations, "observations must be a sequence");
  if (!seq) return NULL;
  const Py_ssize_t B = PySequence_Fast_GET_SIZE(seq);
  Py_buffer bl, bx, bb;
  if (get_buf(o_labels, &bl, B * N * 4, "labels") != 0) { Py_DECREF(seq); return NULL; }
  if (get_buf(o_xyz, &bx, B * N * 3 * 8, "xyz") != 0) { PyBuffer_Release(&bl); Py_DECREF(seq); return NULL; }
  if (get_buf(o_bags, &bb, B * Z * 4, "bags") != 0) { PyBuffer_Release(&bl); PyBuffer_Release(&bx); Py_DECREF(seq); return NULL; }
  int32_t* labels = (int32_t*)bl.buf;
  double* xyz = (double*)bx.buf;
  float* bags = (float*)bb.buf;
  int ok = 1;
  for (Py_ssize_t b = 0; b < B && ok; ++b) {
    PyObject* obs = PySequence_Fast_GET_ITEM(seq, b);
    PyObject* canvas = NULL; PyObject* bag = NULL; PyObject* cseq = NULL; PyObject* bseq = NULL;
    if (!PySequence_Check(obs) || PySequence_Size(obs) != 2) {
      PyErr_Format(PyExc_RuntimeError, "observation %zd is not a (canvas, bag) pair", b); ok = 0; break;
    }
    canvas = PySequence_GetItem(obs, 0);
    bag = PySequence_GetItem(obs, 1);
    cseq = canvas ? PySequence_Fast(canvas, "canvas must be a sequence") : NULL;
    bseq = bag ? PySequence_Fast(bag, "bag must be a sequence") : NULL;
    if (!cseq || !bseq) ok = 0;
    if (ok && PySequence_Fast_GET_SIZE(cseq) != N) {
      PyErr_Format(PyExc_RuntimeError, "canvas %zd holds %zd items, expected %zd", b, PySequence_Fast_GET_SIZE(cseq), N); ok = 0;
    }
    if (ok && PySequence_Fast_GET_SIZE(bseq) != Z) {
      PyErr_Format(PyExc_RuntimeError, "bag %zd holds %zd counts, expected %zd", b, PySequence_Fast_GET_SIZE(bseq), Z); ok = 0;
    }
    for (Py_ssize_t i = 0; i < N && ok; ++i) {
      PyObject* item = PySequence_Fast_GET_ITEM(cseq, i);
      PyObject* iseq = PySequence_Fast(item, "canvas item must be (label, (x, y, z))");
      if (!iseq || PySequence_Fast_GET_SIZE(iseq) != 2) {
        if (iseq) PyErr_Format(PyExc_RuntimeError, "canvas %zd item %zd is not (label, position)", b, i);
        Py_XDECREF(iseq); ok = 0; break;
      }
      long lab = PyLong_AsLong(PySequence_Fast_GET_ITEM(iseq, 0));
      if (lab == -1 && PyErr_Occurred()) { Py_DECREF(iseq); ok = 0; break; }
      labels[b * N + i] = (int32_t)lab;
      PyObject* pseq = PySequence_Fast(PySequence_Fast_GET_ITEM(iseq, 1), "position must be a sequence of 3 numbers");
      if (!pseq || PySequence_Fast_GET_SIZE(pseq) != 3) {
        if (pseq) PyErr_Format(PyExc_RuntimeError, "canvas %zd item %zd: position needs 3 numbers", b, i);
        Py_XDECREF(pseq); Py_DECREF(iseq); ok = 0; break;
      }
      for (int a = 0; a < 3; ++a) {
        double v = PyFloat_AsDouble(PySequence_Fast_GET_ITEM(pseq, a));
        if (v == -1.0 && PyErr_Occurred()) { ok = 0; break; }
        xyz[(b * N + i) * 3 + a] = v;
      }
      Py_DECREF(pseq);
      Py_DECREF(iseq);
    }
    for (Py_ssize_t z = 0; z < Z && ok; ++z) {
      double v = PyFloat_AsDouble(PySequence_Fast_GET_ITEM(bseq, z));
      if (v == -1.0 && PyErr_Occurred()) { ok = 0; break; }
      bags[b * Z + z] = (float)v;
    }
    Py_XDECREF(cseq); Py_XDECREF(bseq); Py_XDECREF(canvas); Py_XDECREF(bag);
  }
  PyBuffer_Release(&bl); PyBuffer_Release(&bx); PyBuffer_Release(&bb);
  Py_DECREF(seq);
  if (!ok) return NULL;
  Py_RETURN_NONE;
}

static PyMethodDef methods[] = {{"flatten", flatten, METH_VARARGS, "flatten observation tuples into arrays"}, {NULL, NULL, 0, NULL}};
static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_mgb_packer", NULL, -1, methods};
PyMODINIT_FUNC PyInit__mgb_packer(void) { return PyModule_Create(&moddef); }
