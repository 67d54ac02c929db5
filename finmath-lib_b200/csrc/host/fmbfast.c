/*
 * _fmbfast — CPython accelerator for the hot path of the Python host binding (finmath-lib_b200/stochastic.py, native.py).
 *
 * The Python classes mirror the Java host side (RandomVariableCuda over the JNI shim); one RandomVariable operation in pure Python
 * costs 4-7 us of interpreter work (type dispatch, filtration-time bookkeeping, ctypes marshalling, two object constructions) - more
 * than the kernel launch it wraps, and a Bermudan valuation issues ~800 of them (profiles/r02_notes.md).  This module implements the
 * COMMON CASES of those operations in C with exactly the semantics of the Python code it short-cuts (which stays in place as the
 * general path and as the specification):
 *     stochastic (device vector) op number,  stochastic op deterministic RandomVariableCuda (add / sub / mult),
 *     stochastic op stochastic,  accrue / discount / addProduct / choose with stochastic operands.
 * Anything else (deterministic receivers, other RandomVariable types, deferred chains, size mismatches) returns NotImplemented and
 * the Python method carries on.  No numerics here: the functions called are fmb_rv_unary / fmb_rv_binary / fmb_rv_ternary /
 * fmb_rv_free of libfinmath_b200.so, bound by address from the ctypes handle (no second copy of the library is loaded).
 *
 * Types: DV (owner of one native handle; base of native.DeviceVector) and RV (time, shard, nGlobal, dv, valueIfNonStochastic; base of
 * stochastic.RandomVariableCuda).
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <structmember.h>
#include <stdint.h>

typedef uint64_t fmb_handle;
typedef int (*unary_fn)(int, fmb_handle, double, fmb_handle*);
typedef int (*binary_fn)(int, fmb_handle, double, fmb_handle, double, fmb_handle*);
typedef int (*ternary_fn)(int, fmb_handle, double, fmb_handle, double, fmb_handle, double, double, fmb_handle*);
typedef int (*free_fn)(fmb_handle);

static unary_fn p_unary;
static binary_fn p_binary;
static ternary_fn p_ternary;
static free_fn p_free;
static PyObject* check_fn;          /* native.check(rc): raises the Python exception for a non-zero return code */
static PyTypeObject* rv_class;      /* stochastic.RandomVariableCuda (results are instances of it) */
static PyTypeObject* dv_class;      /* native.DeviceVector */
static unsigned long long lazy_min_n = ~0ull;   /* vectors at least this long are deferred by the Python layer: not handled here */

/* ---- DV ------------------------------------------------------------------------------------------------------------------ */
typedef struct { PyObject_HEAD unsigned long long h; unsigned long long n; PyObject* weakrefs; } DV;

static int DV_init(DV* self, PyObject* args, PyObject* kw) {
	PyObject *oh, *on;
	if (!PyArg_ParseTuple(args, "OO", &oh, &on)) return -1;
	(void)kw;
	PyObject* ih = PyNumber_Index(oh);                    /* (numpy integers come in as well) */
	if (!ih) return -1;
	PyObject* in = PyNumber_Index(on);
	if (!in) { Py_DECREF(ih); return -1; }
	const unsigned long long h = PyLong_AsUnsignedLongLong(ih), n = PyLong_AsUnsignedLongLong(in);
	Py_DECREF(ih); Py_DECREF(in);
	if (PyErr_Occurred()) return -1;
	self->h = h; self->n = n;
	return 0;
}
static void DV_dealloc(DV* self) {
	if (self->weakrefs) PyObject_ClearWeakRefs((PyObject*)self);
	if (self->h && p_free) { p_free(self->h); self->h = 0; }
	Py_TYPE(self)->tp_free((PyObject*)self);
}
static PyMemberDef DV_members[] = {
	{"h", T_ULONGLONG, offsetof(DV, h), 0, "native handle (0: released)"},
	{"n", T_ULONGLONG, offsetof(DV, n), 0, "number of elements"},
	{NULL, 0, 0, 0, NULL}};
static PyTypeObject DVType = {
	PyVarObject_HEAD_INIT(NULL, 0).tp_name = "_fmbfast.DV", .tp_basicsize = sizeof(DV), .tp_flags = Py_TPFLAGS_DEFAULT | Py_TPFLAGS_BASETYPE,
	.tp_new = PyType_GenericNew, .tp_init = (initproc)DV_init, .tp_dealloc = (destructor)DV_dealloc, .tp_members = DV_members,
	.tp_weaklistoffset = offsetof(DV, weakrefs), .tp_doc = "owner of one device vector handle"};

static PyObject* new_dv(fmb_handle h, unsigned long long n) {
	DV* d = (DV*)dv_class->tp_alloc(dv_class, 0);
	if (!d) { p_free(h); return NULL; }
	d->h = h; d->n = n; d->weakrefs = NULL;
	return (PyObject*)d;
}

/* ---- RV ------------------------------------------------------------------------------------------------------------------ */
typedef struct { PyObject_HEAD double time; double value; PyObject* shard; PyObject* nGlobal; PyObject* dv; } RV;

static PyObject* RV_new(PyTypeObject* type, PyObject* args, PyObject* kw) {
	(void)args; (void)kw;
	RV* self = (RV*)type->tp_alloc(type, 0);
	if (!self) return NULL;
	self->time = 0.0; self->value = Py_NAN;
	Py_INCREF(Py_None); self->shard = Py_None;
	Py_INCREF(Py_None); self->nGlobal = Py_None;
	Py_INCREF(Py_None); self->dv = Py_None;
	return (PyObject*)self;
}
static int RV_traverse(RV* self, visitproc visit, void* arg) { Py_VISIT(self->shard); Py_VISIT(self->nGlobal); Py_VISIT(self->dv); return 0; }
static int RV_clear(RV* self) { Py_CLEAR(self->shard); Py_CLEAR(self->nGlobal); Py_CLEAR(self->dv); return 0; }
static void RV_dealloc(RV* self) {
	PyObject_GC_UnTrack(self);
	RV_clear(self);
	Py_TYPE(self)->tp_free((PyObject*)self);
}
static PyMemberDef RV_members[] = {
	{"time", T_DOUBLE, offsetof(RV, time), 0, "filtration time"},
	{"valueIfNonStochastic", T_DOUBLE, offsetof(RV, value), 0, "value of a deterministic random variable (NaN otherwise)"},
	{"shard", T_OBJECT_EX, offsetof(RV, shard), 0, "ShardContext"},
	{"nGlobal", T_OBJECT_EX, offsetof(RV, nGlobal), 0, "logical number of paths over all shards"},
	{"dv", T_OBJECT_EX, offsetof(RV, dv), 0, "device vector (None: deterministic)"},
	{NULL, 0, 0, 0, NULL}};

/* result object: a RandomVariableCuda with (time, shard and nGlobal of `like`, the new device vector) */
static PyObject* make_result(RV* like, double time, fmb_handle h, unsigned long long n) {
	PyObject* dv = new_dv(h, n);
	if (!dv) return NULL;
	RV* r = (RV*)rv_class->tp_alloc(rv_class, 0);
	if (!r) { Py_DECREF(dv); return NULL; }
	r->time = time; r->value = Py_NAN;
	Py_INCREF(like->shard); r->shard = like->shard;
	Py_INCREF(like->nGlobal); r->nGlobal = like->nGlobal;
	r->dv = dv;
	return (PyObject*)r;
}
static int raise_rc(int rc) {                     /* lets native.check build the exception (ValueError / NotImplementedError / ...) */
	PyObject* r = PyObject_CallFunction(check_fn, "i", rc);
	Py_XDECREF(r);
	if (!PyErr_Occurred()) PyErr_Format(PyExc_RuntimeError, "finmath_b200 error %d", rc);
	return -1;
}
/* the device vector of a stochastic RandomVariableCuda that is ready to use (not a deferred chain), else NULL */
static inline DV* ready_dv(PyObject* o) {
	if (Py_TYPE(o) != rv_class) return NULL;
	PyObject* dv = ((RV*)o)->dv;
	if (Py_TYPE(dv) != dv_class) return NULL;
	return (DV*)dv;
}
static inline int is_det_rv(PyObject* o) { return Py_TYPE(o) == rv_class && ((RV*)o)->dv == Py_None; }
static inline int as_number(PyObject* o, double* v) {
	if (PyFloat_CheckExact(o)) { *v = PyFloat_AS_DOUBLE(o); return 1; }
	if (PyLong_CheckExact(o)) { *v = PyLong_AsDouble(o); if (*v == -1.0 && PyErr_Occurred()) { PyErr_Clear(); return 0; } return 1; }
	return 0;
}
#define NOT_HANDLED do { Py_INCREF(Py_NotImplemented); return Py_NotImplemented; } while (0)

/* fast_unary(op, a): f(x, a) on a stochastic receiver; keeps the receiver's time */
static PyObject* RV_fast_unary(RV* self, PyObject* const* args, Py_ssize_t nargs) {
	if (nargs != 2) { PyErr_SetString(PyExc_TypeError, "fast_unary(op, a)"); return NULL; }
	DV* x = ready_dv((PyObject*)self);
	double a;
	if (!x || x->n >= lazy_min_n || !as_number(args[1], &a)) NOT_HANDLED;
	const int op = (int)PyLong_AsLong(args[0]);
	fmb_handle out = 0;
	const int rc = p_unary(op, x->h, a, &out);
	if (rc) { raise_rc(rc); return NULL; }
	return make_result(self, self->time, out, x->n);
}

/* fast_binary(op_vv, op_vs, other, shortcut): receiver op other.
 *   other a number                       -> unary op_vs (receiver's time)
 *   other a deterministic RV, shortcut 1 -> unary op_vs with its value (receiver's time: "return this.add(rv.doubleValue())")
 *   other a stochastic RV                -> binary op_vv (max of the times) */
static PyObject* RV_fast_binary(RV* self, PyObject* const* args, Py_ssize_t nargs) {
	if (nargs != 4) { PyErr_SetString(PyExc_TypeError, "fast_binary(op_vv, op_vs, other, shortcut)"); return NULL; }
	DV* x = ready_dv((PyObject*)self);
	if (!x || x->n >= lazy_min_n) NOT_HANDLED;
	PyObject* other = args[2];
	double a;
	fmb_handle out = 0;
	int rc;
	if (as_number(other, &a) || (is_det_rv(other) && args[3] == Py_True && (a = ((RV*)other)->value, 1))) {
		const int op = (int)PyLong_AsLong(args[1]);
		if (op < 0) NOT_HANDLED;
		rc = p_unary(op, x->h, a, &out);
		if (rc) { raise_rc(rc); return NULL; }
		return make_result(self, self->time, out, x->n);
	}
	DV* y = ready_dv(other);
	if (!y || y->n != x->n) NOT_HANDLED;
	rc = p_binary((int)PyLong_AsLong(args[0]), x->h, 0.0, y->h, 0.0, &out);
	if (rc) { raise_rc(rc); return NULL; }
	const double t = ((RV*)other)->time;
	return make_result(self, self->time > t ? self->time : t, out, x->n);
}

/* fast_ternary(op, y, z, a): x = receiver.  y / z: stochastic RandomVariableCuda or None (operand absent); all present operands must be
 * stochastic vectors of the receiver's length.  Time = max over the operands. */
static PyObject* RV_fast_ternary(RV* self, PyObject* const* args, Py_ssize_t nargs) {
	if (nargs != 4) { PyErr_SetString(PyExc_TypeError, "fast_ternary(op, y, z, a)"); return NULL; }
	DV* x = ready_dv((PyObject*)self);
	double a;
	if (!x || x->n >= lazy_min_n || !as_number(args[3], &a)) NOT_HANDLED;
	double t = self->time;
	fmb_handle hy = 0, hz = 0;
	if (args[1] != Py_None) {
		DV* y = ready_dv(args[1]);
		if (!y || y->n != x->n) NOT_HANDLED;
		hy = y->h;
		if (((RV*)args[1])->time > t) t = ((RV*)args[1])->time;
	}
	if (args[2] != Py_None) {
		DV* z = ready_dv(args[2]);
		if (!z || z->n != x->n) NOT_HANDLED;
		hz = z->h;
		if (((RV*)args[2])->time > t) t = ((RV*)args[2])->time;
	}
	fmb_handle out = 0;
	const int rc = p_ternary((int)PyLong_AsLong(args[0]), x->h, 0.0, hy, 0.0, hz, 0.0, a, &out);
	if (rc) { raise_rc(rc); return NULL; }
	return make_result(self, t, out, x->n);
}

static PyMethodDef RV_methods[] = {
	{"_fast_unary", (PyCFunction)(void (*)(void))RV_fast_unary, METH_FASTCALL, "f(x, a) on a stochastic receiver, or NotImplemented"},
	{"_fast_binary", (PyCFunction)(void (*)(void))RV_fast_binary, METH_FASTCALL, "receiver op other, or NotImplemented"},
	{"_fast_ternary", (PyCFunction)(void (*)(void))RV_fast_ternary, METH_FASTCALL, "ternary op on stochastic operands, or NotImplemented"},
	{NULL, NULL, 0, NULL}};
static PyTypeObject RVType = {
	PyVarObject_HEAD_INIT(NULL, 0).tp_name = "_fmbfast.RV", .tp_basicsize = sizeof(RV),
	.tp_flags = Py_TPFLAGS_DEFAULT | Py_TPFLAGS_BASETYPE | Py_TPFLAGS_HAVE_GC, .tp_new = RV_new, .tp_dealloc = (destructor)RV_dealloc,
	.tp_traverse = (traverseproc)RV_traverse, .tp_clear = (inquiry)RV_clear, .tp_members = RV_members, .tp_methods = RV_methods,
	.tp_doc = "storage and fast paths of RandomVariableCuda"};

/* ---- module ---------------------------------------------------------------------------------------------------------------- */
/* bind(unary_addr, binary_addr, ternary_addr, free_addr, check, RandomVariableCuda, DeviceVector) */
static PyObject* mod_bind(PyObject* m, PyObject* args) {
	(void)m;
	unsigned long long a0, a1, a2, a3;
	PyObject *chk, *rvc, *dvc;
	if (!PyArg_ParseTuple(args, "KKKKOOO", &a0, &a1, &a2, &a3, &chk, &rvc, &dvc)) return NULL;
	if (!PyType_Check(rvc) || !PyType_IsSubtype((PyTypeObject*)rvc, &RVType) || !PyType_Check(dvc) || !PyType_IsSubtype((PyTypeObject*)dvc, &DVType)) {
		PyErr_SetString(PyExc_TypeError, "bind: RandomVariableCuda / DeviceVector must derive from _fmbfast.RV / _fmbfast.DV");
		return NULL;
	}
	p_unary = (unary_fn)(uintptr_t)a0; p_binary = (binary_fn)(uintptr_t)a1; p_ternary = (ternary_fn)(uintptr_t)a2; p_free = (free_fn)(uintptr_t)a3;
	Py_XDECREF(check_fn); Py_INCREF(chk); check_fn = chk;
	Py_XDECREF((PyObject*)rv_class); Py_INCREF(rvc); rv_class = (PyTypeObject*)rvc;
	Py_XDECREF((PyObject*)dv_class); Py_INCREF(dvc); dv_class = (PyTypeObject*)dvc;
	Py_RETURN_NONE;
}
static PyObject* mod_set_lazy_min_n(PyObject* m, PyObject* arg) {
	(void)m;
	const unsigned long long v = PyLong_AsUnsignedLongLongMask(arg);
	if (PyErr_Occurred()) return NULL;
	lazy_min_n = v;
	Py_RETURN_NONE;
}
static PyObject* mod_bound(PyObject* m, PyObject* noarg) { (void)m; (void)noarg; return PyBool_FromLong(p_unary != NULL && rv_class != NULL && dv_class != NULL); }

static PyMethodDef mod_methods[] = {
	{"bind", mod_bind, METH_VARARGS, "bind the native entry points (addresses from ctypes) and the Python classes"},
	{"set_lazy_min_n", mod_set_lazy_min_n, METH_O, "vectors at least this long are left to the deferred-evaluation layer (2**64-1: never)"},
	{"bound", mod_bound, METH_NOARGS, "True once bind() has run"},
	{NULL, NULL, 0, NULL}};
static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_fmbfast", "fast paths of the finmath_b200 Python host binding", -1, mod_methods, NULL, NULL, NULL, NULL};

PyMODINIT_FUNC PyInit__fmbfast(void) {
	if (PyType_Ready(&DVType) < 0 || PyType_Ready(&RVType) < 0) return NULL;
	PyObject* m = PyModule_Create(&moddef);
	if (!m) return NULL;
	Py_INCREF(&DVType); PyModule_AddObject(m, "DV", (PyObject*)&DVType);
	Py_INCREF(&RVType); PyModule_AddObject(m, "RV", (PyObject*)&RVType);
	return m;
}
