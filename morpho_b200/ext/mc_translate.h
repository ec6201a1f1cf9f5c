/* mc_translate.h — turns a Morpho integrand closure into the postfix program libmorpho_cuda runs (mcu_expr_op).
 *
 * LineIntegral / AreaIntegral call their integrand at every quadrature node by re-entering the VM
 * (integral_integrandfn, functional.c:5146-5170: args = position Matrix, then one interpolated value per Field).
 * The device has no VM, so the function is interpreted ONCE, symbolically: registers hold expression nodes instead of
 * numbers, and the value that reaches OP_RETURN is the integrand as an expression of x[i] and the field components.
 * Handled: straight-line code made of MOV LCT LGL LUP ADD SUB MUL DIV POW LIX CALL INVOKE RETURN (vm.c:924-1620 for
 * their semantics, incl. Int/Float promotion), calls of the real math builtins (functiondefs.c:659-695), calls of other
 * Morpho functions (inlined), element-wise vector arithmetic on the position / matrix-valued quantities, `norm`,
 * `inner`, `sum`, `trace`, `count`, and the element helpers tangent(), normal(), grad(q) (constants of the element).
 * Control flow: comparisons (EQ NEQ LT LE with the VM's tolerance for equal floats), NOT and the branches B BIF BIFF.  A
 * branch on a value known at translation time is simply followed (loops with constant bounds unroll); a branch on a value
 * that depends on the position or a field FORKS the interpretation: both continuations run to their RETURN and the result
 * is SELECT(condition, taken, not taken) — the device evaluates both sides, branch free.
 * Anything else (for-in over collections, lists, printing, writes to globals ...) makes the translation fail, and the caller runs the reference builtin instead.
 * Globals and upvalues are read at translation time; the translation is redone at every evaluation.
 */
#ifndef MC_TRANSLATE_H
#define MC_TRANSLATE_H

#include <math.h>
#include "core.h"
#include "function.h"
#include "closure.h"
#include "upvalue.h"
#include "metafunction.h"
#include "strng.h"

#define MT_MAXNODES 4096
#define MT_MAXCOMP 9
#define MT_MAXDEPTH 6

typedef struct { int op, a, b, d, isint; double c; } mt_node;     /* op: MCU_X_*;  CONST: c (isint: a Morpho Int); SELECT: a ? b : d */
typedef enum { MT_NONE, MT_NUM, MT_VEC, MT_OBJ, MT_GRADLIST, MT_BOOL } mt_kind;     /* MT_BOOL: node[0] evaluates to 1.0 / 0.0 */
/* field >= 0: the symbol still IS the interpolated value of that Field argument (what grad() looks up, functional.c:4644-4652);
 * qbase: its first flattened component.  MT_GRADLIST: grad(q) of a matrix-valued q, a List of dim matrices. */
typedef struct { mt_kind kind; int nrows, ncols; int node[MT_MAXCOMP]; value obj; int field, qbase; } mt_sym;
typedef struct {
    vm *v;
    mt_node nodes[MT_MAXNODES];
    int nnodes, depth, dim;
    int steps, forks;                                    /* budget of interpreted instructions / nested forks, shared by all paths */
    int nfields, fqbase[16], frows[16], fcols[16];      /* the Field arguments: first flattened component and shape */
    value fieldobj[16];                                  /* the Field objects themselves, for grad(Field) (functional.c:4644-4646) */
    const char *why;
} mt_ctx;

static bool mt_fail(mt_ctx *c, const char *why) { if (!c->why) c->why = why; return false; }

static int mt_new(mt_ctx *c, int op, int a, int b, double k, int isint) {
    /* structural sharing: the same operation on the same operands is the same node, so that the two continuations of a
     * forked branch (and `var a = ...; a*a`) share what they have in common and mt_emit computes it once.  SELECT nodes
     * get their third operand after creation and are never shared here. */
    if (op != MCU_X_SELECT)
        for (int i = c->nnodes - 1; i >= 0 && i >= c->nnodes - 512; i--) {
            const mt_node *n = &c->nodes[i];
            if (n->op == op && n->a == a && n->b == b && n->isint == isint && n->op != MCU_X_SELECT &&
                (n->c == k && signbit(n->c) == signbit(k))) return i;
        }
    if (c->nnodes == MT_MAXNODES) { mt_fail(c, "expression too large"); return -1; }
    mt_node *n = &c->nodes[c->nnodes];
    n->op = op; n->a = a; n->b = b; n->d = 0; n->c = k; n->isint = isint;
    return c->nnodes++;
}
static int mt_constf(mt_ctx *c, double x) { return mt_new(c, MCU_X_CONST, 0, 0, x, 0); }
static int mt_consti(mt_ctx *c, int i) { return mt_new(c, MCU_X_CONST, i, 0, (double) i, 1); }
static bool mt_isconst(mt_ctx *c, int n) { return n >= 0 && c->nodes[n].op == MCU_X_CONST; }
static bool mt_isint(mt_ctx *c, int n) { return mt_isconst(c, n) && c->nodes[n].isint; }

static mt_sym mt_num(int node) { mt_sym s; memset(&s, 0, sizeof(s)); s.kind = node >= 0 ? MT_NUM : MT_NONE; s.node[0] = node; s.obj = MORPHO_NIL; s.field = -1; return s; }
static mt_sym mt_none(void) { return mt_num(-1); }

/* ADD SUB MUL DIV POW on two scalars, with the VM's Int/Float rules (vm.c:934-1075) and constant folding */
static int mt_binary(mt_ctx *c, int op, int l, int r) {
    if (l < 0 || r < 0) return -1;
    if (mt_isconst(c, l) && mt_isconst(c, r)) {
        const mt_node *a = &c->nodes[l], *b = &c->nodes[r];
        if (a->isint && b->isint && (op == MCU_X_ADD || op == MCU_X_SUB || op == MCU_X_MUL)) {
            long long x = a->a, y = b->a, z = op == MCU_X_ADD ? x + y : op == MCU_X_SUB ? x - y : x * y;
            if (z > 2147483647LL || z < -2147483648LL) { mt_fail(c, "integer overflow in a constant"); return -1; }
            return mt_consti(c, (int) z);
        }
        switch (op) {
            case MCU_X_ADD: return mt_constf(c, a->c + b->c);
            case MCU_X_SUB: return mt_constf(c, a->c - b->c);
            case MCU_X_MUL: return mt_constf(c, a->c * b->c);
            case MCU_X_DIV: return mt_constf(c, a->c / b->c);
            case MCU_X_POW: return mt_constf(c, pow(a->c, b->c));
            case MCU_X_ATAN2: return mt_constf(c, atan2(a->c, b->c));
        }
    }
    if (op == MCU_X_POW && mt_isint(c, r)) return mt_new(c, MCU_X_POWI, c->nodes[r].a, l, 0.0, 0);   /* pow(x, (double) n) */
    return mt_new(c, op, l, r, 0.0, 0);
}

static double mt_hostfn(int op, double x) {
    switch (op) {
        case MCU_X_SQRT: return sqrt(x);   case MCU_X_SIN: return sin(x);     case MCU_X_COS: return cos(x);
        case MCU_X_EXP: return exp(x);     case MCU_X_LOG: return log(x);     case MCU_X_ABS: return fabs(x);
        case MCU_X_TAN: return tan(x);     case MCU_X_ASIN: return asin(x);   case MCU_X_ACOS: return acos(x);
        case MCU_X_ATAN: return atan(x);   case MCU_X_SINH: return sinh(x);   case MCU_X_COSH: return cosh(x);
        case MCU_X_TANH: return tanh(x);   case MCU_X_FLOOR: return floor(x); case MCU_X_CEIL: return ceil(x);
        case MCU_X_LOG10: return log10(x); case MCU_X_NEG: return -x;
    }
    return NAN;
}

static int mt_unary(mt_ctx *c, int op, int x) {
    if (x < 0) return -1;
    if (mt_isconst(c, x)) return mt_constf(c, mt_hostfn(op, c->nodes[x].c));
    return mt_new(c, op, x, 0, 0.0, 0);      /* operand in a */
}

/* morpho_extendedcomparevalue on numbers (value.c:48-120): 0 equal (floats: within DBL_EPSILON relative), +1 if b > a, else -1 */
static int mt_cmpconst(const mt_node *a, const mt_node *b) {
    if (a->isint && b->isint) return a->a == b->a ? 0 : (b->a > a->a ? 1 : -1);
    const double x = a->c, y = b->c;
    if (x == y) return 0;
    const double diff = fabs(x - y), ax = fabs(x), ay = fabs(y), amax = ax > ay ? ax : ay;
    if (diff == 0.0 || (amax > DBL_MIN && diff / amax <= DBL_EPSILON)) return 0;
    return y > x ? 1 : -1;
}
static mt_sym mt_bool(int node) { mt_sym s; memset(&s, 0, sizeof(s)); s.kind = node >= 0 ? MT_BOOL : MT_NONE; s.node[0] = node; s.obj = MORPHO_NIL; s.field = -1; return s; }
/* EQ NEQ LT LE on two numbers (vm.c:1095-1135) */
static mt_sym mt_compare(mt_ctx *c, int op, const mt_sym *l, const mt_sym *r) {
    if (l->kind != MT_NUM || r->kind != MT_NUM) { mt_fail(c, "comparison of something that is not a number"); return mt_none(); }
    const int a = l->node[0], b = r->node[0];
    if (mt_isconst(c, a) && mt_isconst(c, b)) {
        const int k = mt_cmpconst(&c->nodes[a], &c->nodes[b]);
        const bool t = op == MCU_X_LT ? k > 0 : (op == MCU_X_LE ? k >= 0 : (op == MCU_X_EQ ? k == 0 : k != 0));
        return mt_bool(mt_constf(c, t ? 1.0 : 0.0));
    }
    return mt_bool(mt_new(c, op, a, b, 0.0, 0));
}
/* the truth value a branch sees (morpho_isfalse, value.c:136-138): nil and false are false, everything else is true.
 * Returns 1 / 0 when known at translation time, -1 when it depends on the point (then *node is the 1.0 / 0.0 expression) */
static int mt_truth(mt_ctx *c, const mt_sym *s, int *node) {
    if (s->kind == MT_BOOL) {
        if (mt_isconst(c, s->node[0])) return c->nodes[s->node[0]].c != 0.0 ? 1 : 0;
        *node = s->node[0];
        return -1;
    }
    if (s->kind == MT_NONE) return 0;                                    /* nil */
    if (s->kind == MT_OBJ && MORPHO_ISBOOL(s->obj)) return MORPHO_GETBOOLVALUE(s->obj) ? 1 : 0;
    if (s->kind == MT_OBJ && MORPHO_ISNIL(s->obj)) return 0;
    return 1;
}

/* a runtime value (constant table, global, upvalue, argument) as a symbol */
static mt_sym mt_fromvalue(mt_ctx *c, value val) {
    if (MORPHO_ISFLOAT(val)) return mt_num(mt_constf(c, MORPHO_GETFLOATVALUE(val)));
    if (MORPHO_ISINTEGER(val)) return mt_num(mt_consti(c, MORPHO_GETINTEGERVALUE(val)));
    if (MORPHO_ISBOOL(val)) return mt_bool(mt_constf(c, MORPHO_GETBOOLVALUE(val) ? 1.0 : 0.0));
    mt_sym s = mt_none();
    if (MORPHO_ISMATRIX(val)) {                /* a constant Matrix captured by the closure */
        objectmatrix *m = MORPHO_GETMATRIX(val);
        if (m->nvals == 1 && m->nels >= 1 && m->nels <= MT_MAXCOMP) {
            s.kind = MT_VEC; s.nrows = (int) m->nrows; s.ncols = (int) m->ncols;
            for (int i = 0; i < (int) m->nels; i++) s.node[i] = mt_constf(c, m->elements[i]);
            return s;
        }
    }
    if (MORPHO_ISOBJECT(val)) { s.kind = MT_OBJ; s.obj = val; }
    return s;
}

static const struct { const char *name; int op; } mt_unaryfns[] = {
    {"sqrt", MCU_X_SQRT}, {"sin", MCU_X_SIN}, {"cos", MCU_X_COS}, {"exp", MCU_X_EXP}, {"log", MCU_X_LOG},
    {"abs", MCU_X_ABS}, {"tan", MCU_X_TAN}, {"asin", MCU_X_ASIN}, {"acos", MCU_X_ACOS}, {"arctan", MCU_X_ATAN},
    {"sinh", MCU_X_SINH}, {"cosh", MCU_X_COSH}, {"tanh", MCU_X_TANH}, {"floor", MCU_X_FLOOR}, {"ceil", MCU_X_CEIL},
    {"log10", MCU_X_LOG10}, {NULL, 0}};

/* the name of a builtin (meta)function, only if it really is the builtin of that name */
static const char *mt_builtinname(value fn) {
    value name = MORPHO_NIL;
    if (MORPHO_ISBUILTINFUNCTION(fn)) name = MORPHO_GETBUILTINFUNCTION(fn)->name;
    else if (MORPHO_ISMETAFUNCTION(fn)) {
        objectmetafunction *mf = MORPHO_GETMETAFUNCTION(fn);
        for (unsigned int i = 0; i < mf->fns.count; i++) if (!MORPHO_ISBUILTINFUNCTION(mf->fns.data[i])) return NULL;   /* user overloads */
        name = mf->name;
    }
    if (!MORPHO_ISSTRING(name)) return NULL;
    const char *s = MORPHO_GETCSTRING(name);
    value std = builtin_findfunction(name);     /* shadowed by a user definition? then fn is not what the name says */
    if (MORPHO_ISNIL(std)) return NULL;
    if (MORPHO_GETOBJECT(std) != MORPHO_GETOBJECT(fn)) {
        /* the compiler may have resolved the call to one implementation of the metafunction already, or wrapped the
         * same implementations in a fresh metafunction: accept if an implementation is shared with the standard one */
        if (!MORPHO_ISMETAFUNCTION(std)) return NULL;
        objectmetafunction *sm = MORPHO_GETMETAFUNCTION(std);
        object *probe = MORPHO_ISMETAFUNCTION(fn) ? (MORPHO_GETMETAFUNCTION(fn)->fns.count > 0 ? MORPHO_GETOBJECT(MORPHO_GETMETAFUNCTION(fn)->fns.data[0]) : NULL)
                                                  : MORPHO_GETOBJECT(fn);
        bool found = false;
        for (unsigned int i = 0; i < sm->fns.count && !found; i++) found = probe && MORPHO_GETOBJECT(sm->fns.data[i]) == probe;
        if (!found) return NULL;
    }
    return s;
}

static bool mt_sameshape(const mt_sym *a, const mt_sym *b) { return a->kind == MT_VEC && b->kind == MT_VEC && a->nrows == b->nrows && a->ncols == b->ncols; }

/* reg[a] = reg[b] op reg[c] for scalars and (element-wise) vectors, as Matrix_add/sub/mul/div do */
static bool mt_arith(mt_ctx *c, int op, const mt_sym *lp, const mt_sym *rp, mt_sym *out) {
    const mt_sym lv = *lp, rv = *rp, *l = &lv, *r = &rv;          /* `i = i + 1` writes the register it reads */
    *out = mt_none();
    if (l->kind == MT_NUM && r->kind == MT_NUM) { *out = mt_num(mt_binary(c, op, l->node[0], r->node[0])); return out->kind == MT_NUM || mt_fail(c, "arithmetic"); }
    const int n = (l->kind == MT_VEC ? l : r)->nrows * (l->kind == MT_VEC ? l : r)->ncols;
    if (op == MCU_X_ADD || op == MCU_X_SUB) {
        if (!mt_sameshape(l, r)) return mt_fail(c, "vector arithmetic on operands of different kinds");
        *out = *l; out->field = -1;
        for (int i = 0; i < n; i++) if ((out->node[i] = mt_binary(c, op, l->node[i], r->node[i])) < 0) return false;
        return true;
    }
    if (op == MCU_X_MUL && ((l->kind == MT_VEC && r->kind == MT_NUM) || (l->kind == MT_NUM && r->kind == MT_VEC))) {
        const mt_sym *vec = l->kind == MT_VEC ? l : r, *s = l->kind == MT_VEC ? r : l;
        *out = *vec; out->field = -1;
        for (int i = 0; i < n; i++) if ((out->node[i] = mt_binary(c, MCU_X_MUL, vec->node[i], s->node[0])) < 0) return false;   /* cblas_dscal: x[i]*s */
        return true;
    }
    if (op == MCU_X_DIV && l->kind == MT_VEC && r->kind == MT_NUM) {
        /* Matrix / scalar scales by the reciprocal (matrix.c: Matrix_div → scale by 1/x) */
        int inv = mt_binary(c, MCU_X_DIV, mt_constf(c, 1.0), r->node[0]);
        *out = *l; out->field = -1;
        for (int i = 0; i < n; i++) if ((out->node[i] = mt_binary(c, MCU_X_MUL, l->node[i], inv)) < 0) return false;
        return true;
    }
    return mt_fail(c, "unsupported matrix arithmetic");
}

static bool mt_constindex(mt_ctx *c, const mt_sym *s, int *i) {
    if (s->kind != MT_NUM || !mt_isint(c, s->node[0])) return mt_fail(c, "index is not a constant integer");
    *i = c->nodes[s->node[0]].a;
    return true;
}

static int mt_sumsq(mt_ctx *c, const mt_sym *a, const mt_sym *b) {
    int acc = -1;
    for (int i = 0; i < a->nrows * a->ncols; i++) {
        int t = mt_binary(c, MCU_X_MUL, a->node[i], b->node[i]);
        acc = acc < 0 ? t : mt_binary(c, MCU_X_ADD, acc, t);
        if (acc < 0) return -1;
    }
    return acc;
}

static bool mt_interp(mt_ctx *c, value fn, const mt_sym *args, int nargs, mt_sym *ret);

/* grad(q)[k] for a matrix-valued q: a matrix shaped like q whose entries are d q_c / d x_k */
static bool mt_gradlist_at(mt_ctx *c, const mt_sym *lst, int k, mt_sym *out) {
    if (k < 0 || k >= c->dim) return mt_fail(c, "index out of bounds");
    *out = mt_none();
    out->kind = MT_VEC; out->nrows = lst->nrows; out->ncols = lst->ncols;
    for (int i = 0; i < lst->nrows * lst->ncols; i++) if ((out->node[i] = mt_new(c, MCU_X_GRADQ, 3 * (lst->qbase + i) + k, 0, 0.0, 0)) < 0) return false;
    return true;
}

/* OP_INVOKE on a vector: the Matrix methods an integrand typically uses */
static bool mt_invoke(mt_ctx *c, const char *sel, const mt_sym *obj, const mt_sym *args, int nargs, mt_sym *out) {
    if (obj->kind != MT_VEC) return mt_fail(c, "method call on something that is not a matrix quantity");
    const int n = obj->nrows * obj->ncols;
    if (!strcmp(sel, "norm") && nargs == 0) { *out = mt_num(mt_unary(c, MCU_X_SQRT, mt_sumsq(c, obj, obj))); return out->kind == MT_NUM; }
    if (!strcmp(sel, "inner") && nargs == 1 && mt_sameshape(obj, &args[0])) { *out = mt_num(mt_sumsq(c, obj, &args[0])); return out->kind == MT_NUM; }
    if (!strcmp(sel, "sum") && nargs == 0) {
        int acc = obj->node[0];
        for (int i = 1; i < n; i++) acc = mt_binary(c, MCU_X_ADD, acc, obj->node[i]);
        *out = mt_num(acc); return acc >= 0;
    }
    if (!strcmp(sel, "trace") && nargs == 0 && obj->nrows == obj->ncols) {
        int acc = obj->node[0];
        for (int i = 1; i < obj->nrows; i++) acc = mt_binary(c, MCU_X_ADD, acc, obj->node[i * obj->nrows + i]);
        *out = mt_num(acc); return acc >= 0;
    }
    if (!strcmp(sel, "count") && nargs == 0) { *out = mt_num(mt_consti(c, n)); return true; }
    if (!strcmp(sel, "transpose") && nargs == 0) {
        *out = *obj; out->field = -1; out->nrows = obj->ncols; out->ncols = obj->nrows;
        for (int i = 0; i < obj->nrows; i++) for (int j = 0; j < obj->ncols; j++) out->node[j + i * obj->ncols] = obj->node[i + j * obj->nrows];
        return true;
    }
    return mt_fail(c, "unsupported matrix method");
}

static bool mt_call(mt_ctx *c, const mt_sym *callee, const mt_sym *args, int nargs, mt_sym *out) {
    if (callee->kind != MT_OBJ) return mt_fail(c, "call of a non-function");
    value fn = callee->obj;
    if (MORPHO_ISFUNCTION(fn) || MORPHO_ISCLOSURE(fn)) return mt_interp(c, fn, args, nargs, out);
    const char *name = mt_builtinname(fn);
    if (!name) return mt_fail(c, "call of an unknown builtin");
    /* List(a, b, ...) of numbers and Matrix(list) / Matrix(n): small constant-shape vectors built inside the closure */
    if (!strcmp(name, "List") && nargs >= 1 && nargs <= MT_MAXCOMP) {
        *out = mt_none(); out->kind = MT_VEC; out->nrows = nargs; out->ncols = 1;
        for (int i = 0; i < nargs; i++) { if (args[i].kind != MT_NUM) return mt_fail(c, "List of non-numbers"); out->node[i] = args[i].node[0]; }
        return true;
    }
    if (!strcmp(name, "Matrix") && nargs == 1) {
        if (args[0].kind == MT_VEC && args[0].ncols == 1) { *out = args[0]; out->field = -1; return true; }
        if (args[0].kind == MT_NUM && mt_isint(c, args[0].node[0]) && c->nodes[args[0].node[0]].a >= 1 && c->nodes[args[0].node[0]].a <= MT_MAXCOMP) {
            *out = mt_none(); out->kind = MT_VEC; out->nrows = c->nodes[args[0].node[0]].a; out->ncols = 1;
            for (int i = 0; i < out->nrows; i++) out->node[i] = mt_constf(c, 0.0);
            return true;
        }
        return mt_fail(c, "unsupported Matrix constructor");
    }
    /* element helpers (functional.c:4396-4745): constants of the element, leaves of the expression */
    if (nargs == 0 && (!strcmp(name, "tangent") || !strcmp(name, "normal"))) {
        *out = mt_none(); out->kind = MT_VEC; out->nrows = c->dim; out->ncols = 1;
        for (int i = 0; i < c->dim; i++) out->node[i] = mt_new(c, name[0] == 't' ? MCU_X_TANGENT : MCU_X_NORMAL, i, 0, 0.0, 0);
        return true;
    }
    if (nargs == 1 && !strcmp(name, "grad")) {
        const mt_sym *q = &args[0];
        int fk = q->field;
        if (fk < 0 && q->kind == MT_OBJ)            /* the Field object itself instead of its interpolated value */
            for (int k = 0; k < c->nfields; k++) if (MORPHO_ISOBJECT(c->fieldobj[k]) && MORPHO_GETOBJECT(c->fieldobj[k]) == MORPHO_GETOBJECT(q->obj)) { fk = k; break; }
        if (fk < 0 || fk >= c->nfields) return mt_fail(c, "grad() of something that is not a Field argument");
        if (q->kind == MT_NUM) {
            /* The reference finds the Field of an interpolated NUMBER by comparing values (functional.c:4647-4651) and
             * raises IntgrlAmbgsFld when two scalar quantities coincide at a node: data dependent, so with more than
             * one scalar Field the closure stays on the reference builtin. */
            int nscalar = 0;
            for (int k = 0; k < c->nfields; k++) if (c->frows[k] == 0) nscalar++;
            if (nscalar > 1) return mt_fail(c, "grad() of an interpolated scalar with several scalar Fields (resolved by value in the reference)");
        }
        *out = mt_none();
        if (c->frows[fk] == 0) {                    /* scalar Field: a dim x 1 Matrix (integral_gradalloc) */
            out->kind = MT_VEC; out->nrows = c->dim; out->ncols = 1;
            for (int k = 0; k < c->dim; k++) out->node[k] = mt_new(c, MCU_X_GRADQ, 3 * c->fqbase[fk] + k, 0, 0.0, 0);
        } else {                                    /* Matrix prototype: a List of dim matrices, [k] = d q / d x_k */
            out->kind = MT_GRADLIST; out->nrows = c->frows[fk]; out->ncols = c->fcols[fk]; out->qbase = c->fqbase[fk];
        }
        return true;
    }
    for (int i = 0; i < nargs; i++) if (args[i].kind != MT_NUM) return mt_fail(c, "builtin called with a non-numeric argument");
    if (nargs == 1) for (int k = 0; mt_unaryfns[k].name; k++) if (!strcmp(name, mt_unaryfns[k].name)) {
        *out = mt_num(mt_unary(c, mt_unaryfns[k].op, args[0].node[0]));
        return out->kind == MT_NUM;
    }
    if (nargs == 2 && !strcmp(name, "arctan")) {        /* arctan(x, y) = atan2(y, x) (functiondefs.c:295-300) */
        *out = mt_num(mt_binary(c, MCU_X_ATAN2, args[1].node[0], args[0].node[0]));
        return out->kind == MT_NUM;
    }
    return mt_fail(c, "call of a builtin the device does not have");
}

#define MT_MAXSTEPS 20000
#define MT_MAXFORKS 24
#define MT_NREG (MORPHO_MAXREGISTERS + 8)            /* INVOKE touches reg[a + 2] */

static int mt_select(mt_ctx *c, int cond, int a, int b) {
    if (cond < 0 || a < 0 || b < 0) return -1;
    if (a == b) return a;
    int n = mt_new(c, MCU_X_SELECT, cond, a, 0.0, 0);
    if (n >= 0) c->nodes[n].d = b;
    return n;
}

static bool mt_exec(mt_ctx *c, objectfunction *func, objectclosure *closure, mt_sym *reg, const instruction *pc, mt_sym *ret);

/* a branch whose condition depends on the point: both continuations to their RETURN, merged with SELECT */
static bool mt_fork(mt_ctx *c, objectfunction *func, objectclosure *closure, const mt_sym *reg, int cond,
                    const instruction *pc_true, const instruction *pc_false, mt_sym *ret) {
    if (++c->forks > MT_MAXFORKS) return mt_fail(c, "too many data-dependent branches");
    mt_sym *copy = (mt_sym *) malloc(sizeof(mt_sym) * MT_NREG);
    if (!copy) return mt_fail(c, "allocation");
    mt_sym r1, r2;
    memcpy(copy, reg, sizeof(mt_sym) * MT_NREG);
    bool ok = mt_exec(c, func, closure, copy, pc_true, &r1);
    if (ok) {
        memcpy(copy, reg, sizeof(mt_sym) * MT_NREG);
        ok = mt_exec(c, func, closure, copy, pc_false, &r2);
    }
    free(copy);
    if (!ok) return false;
    if (r1.kind != r2.kind || (r1.kind != MT_NUM && r1.kind != MT_VEC && r1.kind != MT_BOOL) ||
        (r1.kind == MT_VEC && (r1.nrows != r2.nrows || r1.ncols != r2.ncols))) return mt_fail(c, "the branches of an if return different kinds of value");
    *ret = r1;
    ret->field = -1;
    const int ncomp = r1.kind == MT_VEC ? r1.nrows * r1.ncols : 1;
    for (int i = 0; i < ncomp; i++) {
        ret->node[i] = mt_select(c, cond, r1.node[i], r2.node[i]);
        if (ret->node[i] < 0) return false;
    }
    return true;
}

static bool mt_interp(mt_ctx *c, value fn, const mt_sym *args, int nargs, mt_sym *ret) {
    objectclosure *closure = NULL;
    objectfunction *func;
    if (MORPHO_ISCLOSURE(fn)) { closure = MORPHO_GETCLOSURE(fn); func = closure->func; }
    else if (MORPHO_ISFUNCTION(fn)) func = MORPHO_GETFUNCTION(fn);
    else return mt_fail(c, "integrand is not a Morpho function");
    if (func->nargs != nargs || func->varg >= 0 || func->opt.count > 0 || func->nopt > 0) return mt_fail(c, "argument count / optional arguments");
    if (++c->depth > MT_MAXDEPTH) return mt_fail(c, "call depth");
    if (!c->v->instructions) return mt_fail(c, "no program");
    mt_sym *reg = (mt_sym *) malloc(sizeof(mt_sym) * MT_NREG);
    if (!reg) return mt_fail(c, "allocation");
    for (int i = 0; i < MT_NREG; i++) reg[i] = mt_none();
    reg[0] = mt_none(); reg[0].kind = MT_OBJ; reg[0].obj = fn;
    for (int i = 0; i < nargs; i++) reg[i + 1] = args[i];
    for (int i = nargs + 1; i < func->nregs && i <= MORPHO_MAXREGISTERS; i++) reg[i] = mt_num(mt_consti(c, 0));   /* vm.c:733 */
    bool ok = mt_exec(c, func, closure, reg, c->v->instructions + func->entry, ret);
    free(reg);
    c->depth--;
    return ok && !c->why;
}

/* interpret from pc to the RETURN of this function; reg is this path's private register file */
static bool mt_exec(mt_ctx *c, objectfunction *func, objectclosure *closure, mt_sym *reg, const instruction *pc, mt_sym *ret) {
    bool ok = false, done = false;
    for (; !done; pc++) {
        if (++c->steps > MT_MAXSTEPS) { mt_fail(c, "too many instructions (a loop that does not end at translation time?)"); break; }
        const instruction bc = *pc;
        const int a = DECODE_A(bc), b = DECODE_B(bc), cc = DECODE_C(bc);
        switch (DECODE_OP(bc)) {
            case OP_NOP: break;
            case OP_MOV: reg[a] = reg[b]; break;
            case OP_LCT: { int k = DECODE_Bx(bc); if (k >= (int) func->konst.count) { mt_fail(c, "constant index"); done = true; break; } reg[a] = mt_fromvalue(c, func->konst.data[k]); break; }
            case OP_LGL: { int k = DECODE_Bx(bc); if (k >= (int) c->v->globals.count) { mt_fail(c, "global index"); done = true; break; } reg[a] = mt_fromvalue(c, c->v->globals.data[k]); break; }
            case OP_LUP:
                if (!closure || b >= closure->nupvalues || !closure->upvalues[b]) { mt_fail(c, "upvalue"); done = true; break; }
                reg[a] = mt_fromvalue(c, *closure->upvalues[b]->location);
                break;
            case OP_ADD: if (!mt_arith(c, MCU_X_ADD, &reg[b], &reg[cc], &reg[a])) done = true; break;
            case OP_SUB: if (!mt_arith(c, MCU_X_SUB, &reg[b], &reg[cc], &reg[a])) done = true; break;
            case OP_MUL: if (!mt_arith(c, MCU_X_MUL, &reg[b], &reg[cc], &reg[a])) done = true; break;
            case OP_DIV: if (!mt_arith(c, MCU_X_DIV, &reg[b], &reg[cc], &reg[a])) done = true; break;
            case OP_POW:
                if (reg[b].kind != MT_NUM || reg[cc].kind != MT_NUM) { mt_fail(c, "power of a non-number"); done = true; break; }
                reg[a] = mt_num(mt_binary(c, MCU_X_POW, reg[b].node[0], reg[cc].node[0]));
                if (reg[a].kind != MT_NUM) done = true;
                break;
            case OP_LIX: {                          /* reg[b] = reg[a][reg[b] .. reg[c]]  (vm.c:1581-1609) */
                const mt_sym *m = &reg[a];
                int i = 0, j = 0, nidx = cc - b + 1, lin;
                if (m->kind == MT_GRADLIST) {
                    mt_sym g;
                    if (nidx != 1 || !mt_constindex(c, &reg[b], &i) || !mt_gradlist_at(c, m, i, &g)) { mt_fail(c, "indexing grad()"); done = true; break; }
                    reg[b] = g;
                    break;
                }
                if (m->kind != MT_VEC || nidx < 1 || nidx > 2 || !mt_constindex(c, &reg[b], &i) || (nidx == 2 && !mt_constindex(c, &reg[cc], &j))) { mt_fail(c, "indexing"); done = true; break; }
                if (nidx == 1) {                    /* Matrix_enumerate: the i-th stored element */
                    if (i < 0 || i >= m->nrows * m->ncols) { mt_fail(c, "index out of bounds"); done = true; break; }
                    lin = i;
                } else {
                    if (i < 0 || i >= m->nrows || j < 0 || j >= m->ncols) { mt_fail(c, "index out of bounds"); done = true; break; }
                    lin = i + j * m->nrows;
                }
                reg[b] = mt_num(m->node[lin]);
                break;
            }
            case OP_LIXL: {                         /* reg[a] = list reg[b] [ reg[c] ]  (vm.c:1611-1627) */
                int i = 0;
                mt_sym g;
                if (reg[b].kind != MT_GRADLIST || !mt_constindex(c, &reg[cc], &i) || !mt_gradlist_at(c, &reg[b], i, &g)) { mt_fail(c, "list indexing"); done = true; break; }
                reg[a] = g;
                break;
            }
            case OP_CALL: {
                if (cc != 0) { mt_fail(c, "optional arguments in a call"); done = true; break; }
                mt_sym out;
                if (!mt_call(c, &reg[a], &reg[a + 1], b, &out)) { done = true; break; }
                reg[a] = out;
                break;
            }
            case OP_INVOKE: {                       /* selector reg[a], receiver reg[a+1], result reg[a+1] (vm.c:1274-1310) */
                mt_sym out;
                if (cc != 0 || reg[a].kind != MT_OBJ || !MORPHO_ISSTRING(reg[a].obj) ||
                    !mt_invoke(c, MORPHO_GETCSTRING(reg[a].obj), &reg[a + 1], &reg[a + 2], b, &out)) { mt_fail(c, "method call"); done = true; break; }
                reg[a + 1] = out;
                break;
            }
            case OP_EQ: reg[a] = mt_compare(c, MCU_X_EQ, &reg[b], &reg[cc]); if (reg[a].kind != MT_BOOL) done = true; break;
            case OP_NEQ: reg[a] = mt_compare(c, MCU_X_NE, &reg[b], &reg[cc]); if (reg[a].kind != MT_BOOL) done = true; break;
            case OP_LT: reg[a] = mt_compare(c, MCU_X_LT, &reg[b], &reg[cc]); if (reg[a].kind != MT_BOOL) done = true; break;
            case OP_LE: reg[a] = mt_compare(c, MCU_X_LE, &reg[b], &reg[cc]); if (reg[a].kind != MT_BOOL) done = true; break;
            case OP_NOT: {                          /* vm.c:1084-1093: !bool, else "is nil" */
                int node = -1;
                if (reg[b].kind == MT_BOOL) {
                    const int t = mt_truth(c, &reg[b], &node);
                    reg[a] = mt_bool(t >= 0 ? mt_constf(c, t ? 0.0 : 1.0) : mt_new(c, MCU_X_NOT, node, 0, 0.0, 0));
                } else reg[a] = mt_bool(mt_constf(c, reg[b].kind == MT_NONE ? 1.0 : 0.0));
                if (reg[a].kind != MT_BOOL) done = true;
                break;
            }
            case OP_B: pc += DECODE_sBx(bc); break;
            case OP_BIF:
            case OP_BIFF: {
                int node = -1;
                const int t = mt_truth(c, &reg[a], &node);
                const bool jump_if = DECODE_OP(bc) == OP_BIF;
                if (t >= 0) { if ((t == 1) == jump_if) pc += DECODE_sBx(bc); break; }
                /* the condition depends on the point: follow both ways (the loop's pc++ is part of a continuation's start) */
                const instruction *jump = pc + DECODE_sBx(bc) + 1, *fall = pc + 1;
                ok = mt_fork(c, func, closure, reg, node, jump_if ? jump : fall, jump_if ? fall : jump, ret);
                done = true;
                break;
            }
            case OP_RETURN:
                if (a > 0) { *ret = reg[b]; ok = ret->kind == MT_NUM || ret->kind == MT_VEC || ret->kind == MT_BOOL; if (!ok) mt_fail(c, "function returns something that is not a number"); }
                else mt_fail(c, "function returns nil");
                done = true;
                break;
            default:
                mt_fail(c, "an instruction the translator does not handle");
                done = true;
        }
    }
    return ok && !c->why;
}

/* Postfix emission.  A node that is used more than once (and is not a leaf) is computed once: MCU_X_STORE r behind its
 * first evaluation keeps the value in register r of the device's small register file, MCU_X_LOAD r fetches it later.
 * The device evaluates both sides of every SELECT, so "first evaluation" is simply first in program order. */
#define MT_NREGS 32
static bool mt_emit(mt_ctx *c, int root, mcu_expr_op *prog, int *n, int cap) {
    typedef struct { int node, stage; } frame;
    frame *st = (frame *) malloc(sizeof(frame) * (MT_MAXNODES + 8));
    short *uses = (short *) calloc((size_t) c->nnodes + 1, sizeof(short)), *regof = (short *) malloc(sizeof(short) * ((size_t) c->nnodes + 1));
    if (!st || !uses || !regof) { free(st); free(uses); free(regof); return mt_fail(c, "allocation"); }
    for (int i = 0; i < c->nnodes; i++) regof[i] = -1;
#define MT_LEAF(nd) ((nd)->op == MCU_X_CONST || (nd)->op == MCU_X_POS || (nd)->op == MCU_X_FIELD || (nd)->op == MCU_X_TANGENT || \
                     (nd)->op == MCU_X_NORMAL || (nd)->op == MCU_X_GRADQ)
#define MT_BINARY(nd) (((nd)->op >= MCU_X_ADD && (nd)->op <= MCU_X_DIV) || (nd)->op == MCU_X_POW || (nd)->op == MCU_X_ATAN2 || \
                       ((nd)->op >= MCU_X_LT && (nd)->op <= MCU_X_NE) || (nd)->op == MCU_X_SELECT)
    /* pass 1: how often is every node reached from the root (a node's operands are counted once, on its first visit) */
    int sp = 0;
    st[sp++].node = root;
    while (sp > 0) {
        const int id = st[--sp].node;
        if (uses[id]++ > 0) continue;
        const mt_node *nd = &c->nodes[id];
        if (MT_LEAF(nd)) continue;
        if (sp + 3 >= MT_MAXNODES) { free(st); free(uses); free(regof); return mt_fail(c, "expression too deep"); }
        st[sp++].node = nd->op == MCU_X_POWI ? nd->b : nd->a;
        if (MT_BINARY(nd)) st[sp++].node = nd->b;
        if (nd->op == MCU_X_SELECT) st[sp++].node = nd->d;
    }
    /* pass 2: emit */
    int nregs = 0;
    bool ok = true;
    sp = 0;
    st[sp].node = root; st[sp++].stage = 0;
    while (sp > 0 && ok) {
        frame *f = &st[sp - 1];
        const mt_node *nd = &c->nodes[f->node];
        const bool leaf = MT_LEAF(nd), binary = MT_BINARY(nd), ternary = nd->op == MCU_X_SELECT;
        if (sp >= MT_MAXNODES) { ok = mt_fail(c, "expression too deep"); break; }
        if (!leaf && f->stage == 0) {
            if (regof[f->node] >= 0) {                 /* computed before: fetch it */
                if (*n >= cap) { ok = mt_fail(c, "integrand program too long"); break; }
                prog[*n].op = MCU_X_LOAD; prog[*n].a = regof[f->node]; prog[*n].c = 0.0; (*n)++;
                sp--;
                continue;
            }
            f->stage = 1;
            /* POWI keeps its operand in b (a is the exponent); unary ops keep it in a */
            st[sp].node = nd->op == MCU_X_POWI ? nd->b : nd->a; st[sp++].stage = 0;
            continue;
        }
        if (binary && f->stage == 1) { f->stage = 2; st[sp].node = nd->b; st[sp++].stage = 0; continue; }
        if (ternary && f->stage == 2) { f->stage = 3; st[sp].node = nd->d; st[sp++].stage = 0; continue; }
        if (*n >= cap) { ok = mt_fail(c, "integrand program too long"); break; }
        prog[*n].op = nd->op;
        prog[*n].a = ((leaf && nd->op != MCU_X_CONST) || nd->op == MCU_X_POWI) ? nd->a : 0;
        prog[*n].c = nd->op == MCU_X_CONST ? nd->c : 0.0;
        (*n)++;
        if (!leaf && uses[f->node] > 1 && nregs < MT_NREGS) {
            if (*n >= cap) { ok = mt_fail(c, "integrand program too long"); break; }
            regof[f->node] = (short) nregs;
            prog[*n].op = MCU_X_STORE; prog[*n].a = nregs++; prog[*n].c = 0.0; (*n)++;
        }
        sp--;
    }
#undef MT_LEAF
#undef MT_BINARY
    free(st); free(uses); free(regof);
    return ok;
}

/* Entry point.  fieldrows/fieldcols: shape of every Field's Matrix prototype (rows 0 = scalar Field), in argument order. */
static bool mc_translate_integrand(vm *v, value fn, int dim, int nfields, const int *fieldrows, const int *fieldcols,
                                   const value *fieldvals, mcu_expr_op *prog, int cap, int *nops, const char **why) {
    mt_ctx *c = (mt_ctx *) calloc(1, sizeof(mt_ctx));
    if (!c) { *why = "allocation"; return false; }
    c->v = v;
    c->dim = dim;
    mt_sym args[1 + 16];
    bool ok = nfields <= 16;
    if (ok) {
        args[0] = mt_none(); args[0].kind = MT_VEC; args[0].nrows = dim; args[0].ncols = 1;
        for (int i = 0; i < dim; i++) args[0].node[i] = mt_new(c, MCU_X_POS, i, 0, 0.0, 0);
        int q = 0;
        c->nfields = nfields;
        for (int k = 0; k < nfields && ok; k++) {
            mt_sym *s = &args[k + 1];
            *s = mt_none();
            c->fqbase[k] = q; c->frows[k] = fieldrows[k]; c->fcols[k] = fieldcols[k];
            c->fieldobj[k] = fieldvals ? fieldvals[k] : MORPHO_NIL;
            if (fieldrows[k] == 0) {                 /* scalar Field: the quantity is a Float */
                *s = mt_num(mt_new(c, MCU_X_FIELD, q, 0, 0.0, 0));
                s->field = k; s->qbase = q;
                q += 1;
                continue;
            }
            const int n = fieldrows[k] * fieldcols[k];
            if (n < 1 || n > MT_MAXCOMP) { ok = mt_fail(c, "field prototype too large"); break; }
            s->kind = MT_VEC; s->nrows = fieldrows[k]; s->ncols = fieldcols[k];
            s->field = k; s->qbase = q;
            for (int i = 0; i < n; i++) s->node[i] = mt_new(c, MCU_X_FIELD, q + i, 0, 0.0, 0);
            q += n;
        }
    }
    mt_sym ret;
    *nops = 0;
    ok = ok && mt_interp(c, fn, args, nfields + 1, &ret);
    if (ok && ret.kind != MT_NUM) ok = mt_fail(c, "integrand does not return a number");
    ok = ok && mt_emit(c, ret.node[0], prog, nops, cap);
    *why = c->why ? c->why : "";
    free(c);
    return ok;
}

/* ScalarPotential: fn(x, y[, z]) of the coordinates (scalarpotential_integrand, functional.c:2266-2280).  A function
 * that returns a number gives one program; one that returns a Matrix of dim entries (the gradient function,
 * scalarpotential_gradient 2283-2303) gives dim programs, concatenated in prog with their lengths in nops[]. */
static bool mc_translate_pointfn(vm *v, value fn, int dim, mcu_expr_op *prog, int cap, int *nops, int *nprog, const char **why) {
    mt_ctx *c = (mt_ctx *) calloc(1, sizeof(mt_ctx));
    if (!c) { *why = "allocation"; return false; }
    c->v = v;
    c->dim = dim;
    mt_sym args[3], ret;
    for (int i = 0; i < dim && i < 3; i++) args[i] = mt_num(mt_new(c, MCU_X_POS, i, 0, 0.0, 0));
    bool ok = dim >= 1 && dim <= 3 && mt_interp(c, fn, args, dim, &ret);
    if (ok && ret.kind != MT_NUM && ret.kind != MT_VEC) ok = mt_fail(c, "function does not return a number or a Matrix");
    *nprog = 0;
    int total = 0;
    if (ok) {
        const int n = ret.kind == MT_NUM ? 1 : ret.nrows * ret.ncols;
        if (ret.kind == MT_VEC && n != dim) ok = mt_fail(c, "gradient function returns a Matrix of the wrong size");
        for (int k = 0; k < n && ok; k++) {
            int len = 0;
            ok = mt_emit(c, ret.node[k], prog + total, &len, cap - total);
            nops[k] = len; total += len;
        }
        if (ok) *nprog = n;
    }
    *why = c->why ? c->why : "";
    free(c);
    return ok;
}

#endif
