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
 * Anything else (deterministic receivers, other RandomVariable types, size mismatches) returns NotImplemented and
 * the Python method carries on.  No numerics here: the functions called are fmb_rv_unary / fmb_rv_binary / fmb_rv_ternary /
 * fmb_rv_free of libfinmath_b200.so, bound by address from the ctypes handle (no second copy of the library is loaded).
 *
 * Deferred element-wise arithmetic (LV): for vectors of at least lazy_min_n elements an operation is RECORDED, not launched.  When the
 * result is consumed by the next element-wise operation the chain grows; when anything else needs it (a reduction, a kernel argument, a
 * download: every access to `.h`) the whole chain is evaluated in ONE pass by fmb_rv_eval_chain - same device functions, order and
 * roundings as the one-operation kernels (bit-identical; tests/test_gpu_rv.py compares both modes).  A chain is a linear accumulator
 * program: start leaf, then up to 16 instructions (kind, op, position of the accumulator among the operands, the other operands: leaf
 * vectors or scalars); at most 8 distinct leaf vectors and 24 scalars.  Only the FIRST consumer extends a pending chain; a second
 * consumer evaluates it and uses the result as a leaf.
 *
 * Types: DV (owner of one native handle; base of native.DeviceVector), LV (a recorded chain, duck-types DV: h, n; base of
 * native.LazyVector) and RV (time, shard, nGlobal, dv, valueIfNonStochastic; base of stochastic.RandomVariableCuda).
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
typedef int (*chain_fn)(int, const unsigned char*, int, const fmb_handle*, int, const double*, int, fmb_handle*);

static unary_fn p_unary;
static binary_fn p_binary;
static ternary_fn p_ternary;
static free_fn p_free;
static chain_fn p_chain;
static PyObject* check_fn;          /* native.check(rc): raises the Python exception for a non-zero return code */
static PyTypeObject* rv_class;      /* stochastic.RandomVariableCuda (results are instances of it) */
static PyTypeObject* dv_class;      /* native.DeviceVector */
static PyTypeObject* lv_class;      /* native.LazyVector */
static unsigned long long lazy_min_n = ~0ull;   /* vectors at least this long are recorded (LV) instead of launched */

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


/* ---- LV: a recorded chain ---------------------------------------------------------------------------------------------------- */
#define CHAIN_MAX_INSTR 16
#define CHAIN_MAX_LEAVES 8
#define CHAIN_MAX_SCALARS 24
typedef struct {
	unsigned char kind, op, pos, nother;   /* kind 0 unary, 1 binary, 2 ternary; pos: position of the accumulator among the operands */
	double a;                              /* the operation's own double argument (kinds 0 and 2) */
	PyObject* other[2];                    /* the other operands in positional order: DV / LV (owned), or NULL = the scalar in sother */
	double sother[2];
} Instr;
typedef struct {
	PyObject_HEAD
	unsigned long long h;                  /* 0 while pending */
	unsigned long long n;
	PyObject* start;                       /* the leaf the accumulator starts from (DV or LV, owned) */
	Instr* prog;
	int nprog, uses, nvec, nsca;           /* nvec / nsca: upper bounds of the distinct leaf vectors / scalars the chain needs */
	PyObject* weakrefs;
} LV;

static int raise_rc(int rc);
static int lv_materialize(LV* v);

static void lv_release_program(LV* v) {
	if (v->prog) {
		for (int i = 0; i < v->nprog; i++) { Py_XDECREF(v->prog[i].other[0]); Py_XDECREF(v->prog[i].other[1]); }
		PyMem_Free(v->prog);
		v->prog = NULL;
	}
	v->nprog = 0;
	Py_CLEAR(v->start);
}
static void LV_dealloc(LV* self) {
	if (self->weakrefs) PyObject_ClearWeakRefs((PyObject*)self);
	lv_release_program(self);
	if (self->h && p_free) { p_free(self->h); self->h = 0; }
	Py_TYPE(self)->tp_free((PyObject*)self);
}
static PyObject* LV_get_h(LV* self, void* c) {
	(void)c;
	if (!self->h && lv_materialize(self) < 0) return NULL;
	return PyLong_FromUnsignedLongLong(self->h);
}
static int LV_set_h(LV* self, PyObject* v, void* c) {       /* (tests hand out fake handles and clear them again) */
	(void)c;
	if (!v) { PyErr_SetString(PyExc_TypeError, "cannot delete h"); return -1; }
	const unsigned long long h = PyLong_AsUnsignedLongLong(v);
	if (PyErr_Occurred()) return -1;
	self->h = h;
	return 0;
}
static PyObject* LV_pending(LV* self, PyObject* noarg) { (void)noarg; return PyBool_FromLong(self->h == 0); }
static PyObject* LV_length(LV* self, PyObject* noarg) { (void)noarg; return PyLong_FromLong(self->h ? 0 : self->nprog); }
static PyGetSetDef LV_getset[] = {{"h", (getter)LV_get_h, (setter)LV_set_h, "native handle: evaluates the recorded chain on first access", NULL}, {NULL, NULL, NULL, NULL, NULL}};
static PyMemberDef LV_members[] = {{"n", T_ULONGLONG, offsetof(LV, n), READONLY, "number of elements"}, {NULL, 0, 0, 0, NULL}};
static PyMethodDef LV_methods[] = {
	{"pending", (PyCFunction)LV_pending, METH_NOARGS, "True while the chain has not been evaluated"},
	{"chain_length", (PyCFunction)LV_length, METH_NOARGS, "instructions recorded and not yet evaluated"},
	{NULL, NULL, 0, NULL}};
static PyTypeObject LVType = {
	PyVarObject_HEAD_INIT(NULL, 0).tp_name = "_fmbfast.LV", .tp_basicsize = sizeof(LV), .tp_flags = Py_TPFLAGS_DEFAULT | Py_TPFLAGS_BASETYPE,
	.tp_dealloc = (destructor)LV_dealloc, .tp_members = LV_members, .tp_methods = LV_methods, .tp_getset = LV_getset,
	.tp_weaklistoffset = offsetof(LV, weakrefs), .tp_doc = "element-wise operations recorded for one fused evaluation"};

static inline int is_dv(PyObject* o) { return Py_TYPE(o) == dv_class; }
static inline int is_lv(PyObject* o) { return Py_TYPE(o) == lv_class; }
static inline unsigned long long vec_n(PyObject* o) { return is_dv(o) ? ((DV*)o)->n : ((LV*)o)->n; }
/* handle of a DV or LV (evaluating a pending chain); 0 with an exception set on failure */
static fmb_handle vec_handle(PyObject* o) {
	if (is_dv(o)) return ((DV*)o)->h;
	LV* v = (LV*)o;
	if (!v->h && lv_materialize(v) < 0) return 0;
	return v->h;
}

static int lv_materialize(LV* v) {
	fmb_handle out = 0;
	int rc;
	if (!v->start || v->nprog < 1 || !v->prog) { PyErr_SetString(PyExc_ValueError, "finmath_b200: an empty LazyVector (they are created by operations, not directly)"); return -1; }
	if (v->nprog == 1) {                                   /* one operation: the specialised kernel */
		const Instr* in = &v->prog[0];
		fmb_handle hs[3] = {0, 0, 0};
		double sc[3] = {0.0, 0.0, 0.0};
		const int nops = in->nother + 1;
		for (int k = 0, j = 0; k < nops; k++) {
			if (k == in->pos) { hs[k] = vec_handle(v->start); if (!hs[k]) return -1; }
			else { if (in->other[j]) { hs[k] = vec_handle(in->other[j]); if (!hs[k]) return -1; } else sc[k] = in->sother[j]; j++; }
		}
		if (in->kind == 0) rc = p_unary(in->op, hs[0], in->a, &out);
		else if (in->kind == 1) rc = p_binary(in->op, hs[0], sc[0], hs[1], sc[1], &out);
		else rc = p_ternary(in->op, hs[0], sc[0], hs[1], sc[1], hs[2], sc[2], in->a, &out);
	} else {
		PyObject* leaves[CHAIN_MAX_LEAVES + 2];
		fmb_handle hs[CHAIN_MAX_LEAVES + 2];
		double scalars[CHAIN_MAX_SCALARS + 3];
		unsigned char code[8 * CHAIN_MAX_INSTR];
		int nl = 0, ns = 0;
		leaves[nl++] = v->start;
		for (int i = 0; i < v->nprog; i++) {
			const Instr* in = &v->prog[i];
			unsigned char refs[3] = {0, 0, 0};
			for (int k = 0; k < in->nother; k++) {
				if (!in->other[k]) {
					const double o = in->sother[k];
					const int mergeable = (o == o) && (o != 0.0);       /* NaN and signed zeros are never merged (0.0 == -0.0) */
					int idx = -1;
					if (mergeable) for (int q = 0; q < ns; q++) if (scalars[q] == o) { idx = q; break; }
					if (idx < 0) { if (ns >= CHAIN_MAX_SCALARS) goto overflow; idx = ns; scalars[ns++] = o; }
					refs[k] = (unsigned char)(128 | idx);
				} else {
					int idx = -1;
					for (int q = 0; q < nl; q++) if (leaves[q] == in->other[k]) { idx = q; break; }
					if (idx < 0) { if (nl >= CHAIN_MAX_LEAVES) goto overflow; idx = nl; leaves[nl++] = in->other[k]; }
					refs[k] = (unsigned char)idx;
				}
			}
			if (in->kind != 1) {                               /* the operation's own double argument */
				const double o = in->a;
				const int mergeable = (o == o) && (o != 0.0);
				int idx = -1;
				if (mergeable) for (int q = 0; q < ns; q++) if (scalars[q] == o) { idx = q; break; }
				if (idx < 0) { if (ns >= CHAIN_MAX_SCALARS) goto overflow; idx = ns; scalars[ns++] = o; }
				refs[in->kind == 2 ? 2 : 0] = (unsigned char)(128 | idx);
			}
			unsigned char* c = code + 8 * i;
			c[0] = in->kind; c[1] = in->op; c[2] = in->pos; c[3] = refs[0]; c[4] = refs[1]; c[5] = refs[2]; c[6] = 0; c[7] = 0;
		}
		for (int q = 0; q < nl; q++) { hs[q] = vec_handle(leaves[q]); if (!hs[q]) return -1; }
		if (ns == 0) scalars[0] = 0.0;
		rc = p_chain(v->nprog, code, 0, hs, nl, scalars, ns, &out);
	}
	if (rc) { raise_rc(rc); return -1; }
	v->h = out;
	lv_release_program(v);
	return 0;
overflow:
	PyErr_SetString(PyExc_RuntimeError, "finmath_b200: deferred chain exceeds its leaf / scalar budget");   /* (the bounds kept by lazy_op exclude this) */
	return -1;
}

/* Record `kind/op` on the operands (ops[k]: DV / LV, or NULL = the scalar sc[k]); nops = 1, 2 or 3.  Returns a new LV (or NULL with an
 * exception).  The pending chain of the FIRST operand that is an unconsumed LV with room left is extended; other pending operands are
 * evaluated and become leaves. */
static PyObject* lazy_record(int kind, int op, PyObject* const* ops, const double* sc, int nops, double a) {
	long long n = -1;
	int host = -1, nvec = 0, nsca = 0;
	for (int i = 0; i < nops; i++) {
		PyObject* o = ops[i];
		if (!o) { nsca++; continue; }
		const unsigned long long on = vec_n(o);
		if (n < 0) n = (long long)on;
		else if ((unsigned long long)n != on) { PyErr_Format(PyExc_ValueError, "finmath_b200: operand sizes differ (%llu vs %lld)", on, n); return NULL; }
		nvec++;
		if (is_lv(o) && ((LV*)o)->h == 0) {
			LV* l = (LV*)o;
			if (host < 0 && l->uses == 0 && l->nprog < CHAIN_MAX_INSTR && l->nvec + 2 <= CHAIN_MAX_LEAVES && l->nsca + 3 <= CHAIN_MAX_SCALARS) host = i;
			else if (lv_materialize(l) < 0) return NULL;       /* second consumer, or the chain is full: evaluate it, use it as a leaf */
		}
	}
	if (kind != 1) nsca++;
	if (n < 0) { PyErr_SetString(PyExc_ValueError, "finmath_b200: no vector operand"); return NULL; }
	LV* base = NULL;
	if (host >= 0 && ((LV*)ops[host])->h == 0) base = (LV*)ops[host];   /* (not pending any more if the same object sits in two positions) */
	int start = 0;
	if (!base) { while (!ops[start]) start++; }
	const int pos = base ? host : start;
	const int nold = base ? base->nprog : 0;
	LV* r = (LV*)lv_class->tp_alloc(lv_class, 0);
	if (!r) return NULL;
	r->h = 0; r->n = (unsigned long long)n; r->uses = 0; r->weakrefs = NULL; r->start = NULL; r->nprog = 0;
	r->prog = (Instr*)PyMem_Malloc(sizeof(Instr) * (size_t)(nold + 1));
	if (!r->prog) { Py_DECREF(r); PyErr_NoMemory(); return NULL; }
	for (int i = 0; i < nold; i++) {
		r->prog[i] = base->prog[i];
		Py_XINCREF(r->prog[i].other[0]); Py_XINCREF(r->prog[i].other[1]);
	}
	Instr* in = &r->prog[nold];
	in->kind = (unsigned char)kind; in->op = (unsigned char)op; in->pos = (unsigned char)pos; in->a = a;
	in->other[0] = in->other[1] = NULL; in->sother[0] = in->sother[1] = 0.0;
	int j = 0;
	for (int k = 0; k < nops; k++) {
		if (k == pos) continue;
		if (ops[k]) { Py_INCREF(ops[k]); in->other[j] = ops[k]; } else in->sother[j] = sc[k];
		j++;
	}
	in->nother = (unsigned char)j;
	r->nprog = nold + 1;
	if (base) {
		base->uses = 1;
		Py_INCREF(base->start); r->start = base->start;
		r->nvec = base->nvec + nvec - 1; r->nsca = base->nsca + nsca;
	} else {
		Py_INCREF(ops[start]); r->start = ops[start];
		r->nvec = nvec; r->nsca = nsca;
	}
	return (PyObject*)r;
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

/* result object: a RandomVariableCuda with (time, shard and nGlobal of `like`, the given vector: reference stolen) */
static PyObject* make_result_vec(RV* like, double time, PyObject* vec) {
	if (!vec) return NULL;
	RV* r = (RV*)rv_class->tp_alloc(rv_class, 0);
	if (!r) { Py_DECREF(vec); return NULL; }
	r->time = time; r->value = Py_NAN;
	Py_INCREF(like->shard); r->shard = like->shard;
	Py_INCREF(like->nGlobal); r->nGlobal = like->nGlobal;
	r->dv = vec;
	return (PyObject*)r;
}
static PyObject* make_result(RV* like, double time, fmb_handle h, unsigned long long n) { return make_result_vec(like, time, new_dv(h, n)); }
static int raise_rc(int rc) {                     /* lets native.check build the exception (ValueError / NotImplementedError / ...) */
	PyObject* r = PyObject_CallFunction(check_fn, "i", rc);
	Py_XDECREF(r);
	if (!PyErr_Occurred()) PyErr_Format(PyExc_RuntimeError, "finmath_b200 error %d", rc);
	return -1;
}
/* the device vector (DV, or LV: recorded chain) of a stochastic RandomVariableCuda, else NULL */
static inline PyObject* vec_of(PyObject* o) {
	if (Py_TYPE(o) != rv_class) return NULL;
	PyObject* dv = ((RV*)o)->dv;
	return (is_dv(dv) || is_lv(dv)) ? dv : NULL;
}
static inline int is_det_rv(PyObject* o) { return Py_TYPE(o) == rv_class && ((RV*)o)->dv == Py_None; }
static inline int as_number(PyObject* o, double* v) {
	if (PyFloat_CheckExact(o)) { *v = PyFloat_AS_DOUBLE(o); return 1; }
	if (PyLong_CheckExact(o)) { *v = PyLong_AsDouble(o); if (*v == -1.0 && PyErr_Occurred()) { PyErr_Clear(); return 0; } return 1; }
	return 0;
}
#define NOT_HANDLED do { Py_INCREF(Py_NotImplemented); return Py_NotImplemented; } while (0)

/* one element-wise operation on ready operands (ops[k] NULL: scalar sc[k]): recorded if the vectors are long enough, else launched */
static PyObject* apply_op(RV* like, double time, int kind, int op, PyObject* const* ops, const double* sc, int nops, double a, unsigned long long n) {
	if (n >= lazy_min_n) return make_result_vec(like, time, lazy_record(kind, op, ops, sc, nops, a));
	fmb_handle hs[3] = {0, 0, 0}, out = 0;
	for (int k = 0; k < nops; k++) if (ops[k]) { hs[k] = vec_handle(ops[k]); if (!hs[k]) return NULL; }
	int rc;
	if (kind == 0) rc = p_unary(op, hs[0], a, &out);
	else if (kind == 1) rc = p_binary(op, hs[0], sc[0], hs[1], sc[1], &out);
	else rc = p_ternary(op, hs[0], sc[0], hs[1], sc[1], hs[2], sc[2], a, &out);
	if (rc) { raise_rc(rc); return NULL; }
	return make_result(like, time, out, n);
}

/* fast_unary(op, a): f(x, a) on a stochastic receiver; keeps the receiver's time */
static PyObject* RV_fast_unary(RV* self, PyObject* const* args, Py_ssize_t nargs) {
	if (nargs != 2) { PyErr_SetString(PyExc_TypeError, "fast_unary(op, a)"); return NULL; }
	PyObject* x = vec_of((PyObject*)self);
	double a;
	if (!x || !as_number(args[1], &a)) NOT_HANDLED;
	const double sc[1] = {0.0};
	return apply_op(self, self->time, 0, (int)PyLong_AsLong(args[0]), &x, sc, 1, a, vec_n(x));
}

/* fast_binary(op_vv, op_vs, other, shortcut): receiver op other.
 *   other a number                       -> unary op_vs (receiver's time)
 *   other a deterministic RV, shortcut 1 -> unary op_vs with its value (receiver's time: "return this.add(rv.doubleValue())")
 *   other a stochastic RV                -> binary op_vv (max of the times) */
static PyObject* RV_fast_binary(RV* self, PyObject* const* args, Py_ssize_t nargs) {
	if (nargs != 4) { PyErr_SetString(PyExc_TypeError, "fast_binary(op_vv, op_vs, other, shortcut)"); return NULL; }
	PyObject* x = vec_of((PyObject*)self);
	if (!x) NOT_HANDLED;
	PyObject* other = args[2];
	double a;
	const double sc[2] = {0.0, 0.0};
	if (as_number(other, &a) || (is_det_rv(other) && args[3] == Py_True && (a = ((RV*)other)->value, 1))) {
		const int op = (int)PyLong_AsLong(args[1]);
		if (op < 0) NOT_HANDLED;
		return apply_op(self, self->time, 0, op, &x, sc, 1, a, vec_n(x));
	}
	PyObject* y = vec_of(other);
	if (!y || vec_n(y) != vec_n(x)) NOT_HANDLED;
	PyObject* ops[2] = {x, y};
	const double t = ((RV*)other)->time;
	return apply_op(self, self->time > t ? self->time : t, 1, (int)PyLong_AsLong(args[0]), ops, sc, 2, 0.0, vec_n(x));
}

/* fast_ternary(op, y, z, a): x = receiver.  y / z: stochastic RandomVariableCuda or None (operand absent); all present operands must be
 * stochastic vectors of the receiver's length.  Time = max over the operands. */
static PyObject* RV_fast_ternary(RV* self, PyObject* const* args, Py_ssize_t nargs) {
	if (nargs != 4) { PyErr_SetString(PyExc_TypeError, "fast_ternary(op, y, z, a)"); return NULL; }
	PyObject* x = vec_of((PyObject*)self);
	double a;
	if (!x || !as_number(args[3], &a)) NOT_HANDLED;
	double t = self->time;
	PyObject* ops[3] = {x, NULL, NULL};
	const double sc[3] = {0.0, 0.0, 0.0};
	for (int k = 1; k <= 2; k++) {
		if (args[k] == Py_None) continue;
		PyObject* v = vec_of(args[k]);
		if (!v || vec_n(v) != vec_n(x)) NOT_HANDLED;
		ops[k] = v;
		if (((RV*)args[k])->time > t) t = ((RV*)args[k])->time;
	}
	return apply_op(self, t, 2, (int)PyLong_AsLong(args[0]), ops, sc, 3, a, vec_n(x));
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
/* bind(unary_addr, binary_addr, ternary_addr, free_addr, eval_chain_addr, check, RandomVariableCuda, DeviceVector, LazyVector) */
static PyObject* mod_bind(PyObject* m, PyObject* args) {
	(void)m;
	unsigned long long a0, a1, a2, a3, a4;
	PyObject *chk, *rvc, *dvc, *lvc;
	if (!PyArg_ParseTuple(args, "KKKKKOOOO", &a0, &a1, &a2, &a3, &a4, &chk, &rvc, &dvc, &lvc)) return NULL;
	if (!PyType_Check(rvc) || !PyType_IsSubtype((PyTypeObject*)rvc, &RVType) || !PyType_Check(dvc) || !PyType_IsSubtype((PyTypeObject*)dvc, &DVType) ||
	    !PyType_Check(lvc) || !PyType_IsSubtype((PyTypeObject*)lvc, &LVType)) {
		PyErr_SetString(PyExc_TypeError, "bind: RandomVariableCuda / DeviceVector / LazyVector must derive from _fmbfast.RV / DV / LV");
		return NULL;
	}
	p_unary = (unary_fn)(uintptr_t)a0; p_binary = (binary_fn)(uintptr_t)a1; p_ternary = (ternary_fn)(uintptr_t)a2; p_free = (free_fn)(uintptr_t)a3;
	p_chain = (chain_fn)(uintptr_t)a4;
	Py_XDECREF(check_fn); Py_INCREF(chk); check_fn = chk;
	Py_XDECREF((PyObject*)rv_class); Py_INCREF(rvc); rv_class = (PyTypeObject*)rvc;
	Py_XDECREF((PyObject*)dv_class); Py_INCREF(dvc); dv_class = (PyTypeObject*)dvc;
	Py_XDECREF((PyObject*)lv_class); Py_INCREF(lvc); lv_class = (PyTypeObject*)lvc;
	Py_RETURN_NONE;
}
/* lazy_op(kind, op, operands, a): operands = tuple of DeviceVector / LazyVector / float (scalar broadcast) in positional order */
static PyObject* mod_lazy_op(PyObject* m, PyObject* args) {
	(void)m;
	int kind, op;
	PyObject* operands;
	double a;
	if (!PyArg_ParseTuple(args, "iiO!d", &kind, &op, &PyTuple_Type, &operands, &a)) return NULL;
	const Py_ssize_t nops = PyTuple_GET_SIZE(operands);
	if (!lv_class || kind < 0 || kind > 2 || nops != kind + 1) { PyErr_SetString(PyExc_ValueError, "lazy_op(kind, op, operands, a)"); return NULL; }
	PyObject* ops[3] = {NULL, NULL, NULL};
	double sc[3] = {0.0, 0.0, 0.0};
	for (Py_ssize_t k = 0; k < nops; k++) {
		PyObject* o = PyTuple_GET_ITEM(operands, k);
		if (is_dv(o) || is_lv(o)) ops[k] = o;
		else if (!as_number(o, &sc[k])) { PyErr_SetString(PyExc_TypeError, "lazy_op: operands are DeviceVector / LazyVector / float"); return NULL; }
	}
	return lazy_record(kind, op, ops, sc, (int)nops, a);
}
static PyObject* mod_set_lazy_min_n(PyObject* m, PyObject* arg) {
	(void)m;
	const unsigned long long v = PyLong_AsUnsignedLongLongMask(arg);
	if (PyErr_Occurred()) return NULL;
	lazy_min_n = v;
	Py_RETURN_NONE;
}
static PyObject* mod_bound(PyObject* m, PyObject* noarg) { (void)m; (void)noarg; return PyBool_FromLong(p_unary != NULL && rv_class != NULL && dv_class != NULL && lv_class != NULL); }

static PyMethodDef mod_methods[] = {
	{"bind", mod_bind, METH_VARARGS, "bind the native entry points (addresses from ctypes) and the Python classes"},
	{"lazy_op", mod_lazy_op, METH_VARARGS, "record one element-wise operation: a LazyVector"},
	{"set_lazy_min_n", mod_set_lazy_min_n, METH_O, "vectors at least this long are left to the deferred-evaluation layer (2**64-1: never)"},
	{"bound", mod_bound, METH_NOARGS, "True once bind() has run"},
	{NULL, NULL, 0, NULL}};
static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_fmbfast", "fast paths of the finmath_b200 Python host binding", -1, mod_methods, NULL, NULL, NULL, NULL};

PyMODINIT_FUNC PyInit__fmbfast(void) {
	LVType.tp_new = PyType_GenericNew;
	if (PyType_Ready(&DVType) < 0 || PyType_Ready(&LVType) < 0 || PyType_Ready(&RVType) < 0) return NULL;
	PyObject* m = PyModule_Create(&moddef);
	if (!m) return NULL;
	Py_INCREF(&DVType); PyModule_AddObject(m, "DV", (PyObject*)&DVType);
	Py_INCREF(&LVType); PyModule_AddObject(m, "LV", (PyObject*)&LVType);
	Py_INCREF(&RVType); PyModule_AddObject(m, "RV", (PyObject*)&RVType);
	return m;
}
