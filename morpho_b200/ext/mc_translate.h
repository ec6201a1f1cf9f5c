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
 * Anything else (branches, loops, lists, printing, writes to globals ...) makes the translation fail, and the caller runs the reference builtin instead.
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

typedef struct { int op, a, b, isint; double c; } mt_node;        /* op: MCU_X_*;  CONST: c (isint: a Morpho Int) */
typedef enum { MT_NONE, MT_NUM, MT_VEC, MT_OBJ, MT_GRADLIST } mt_kind;
/* field >= 0: the symbol still IS the interpolated value of that Field argument (what grad() looks up, functional.c:4644-4652);
 * qbase: its first flattened component.  MT_GRADLIST: grad(q) of a matrix-valued q, a List of dim matrices. */
typedef struct { mt_kind kind; int nrows, ncols; int node[MT_MAXCOMP]; value obj; int field, qbase; } mt_sym;
typedef struct {
    vm *v;
    mt_node nodes[MT_MAXNODES];
    int nnodes, depth, dim;
    int nfields, fqbase[16], frows[16], fcols[16];      /* the Field arguments: first flattened component and shape */
    value fieldobj[16];                                  /* the Field objects themselves, for grad(Field) (functional.c:4644-4646) */
    const char *why;
} mt_ctx;

static bool mt_fail(mt_ctx *c, const char *why) { if (!c->why) c->why = why; return false; }

static int mt_new(mt_ctx *c, int op, int a, int b, double k, int isint) {
    if (c->nnodes == MT_MAXNODES) { mt_fail(c, "expression too large"); return -1; }
    mt_node *n = &c->nodes[c->nnodes];
    n->op = op; n->a = a; n->b = b; n->c = k; n->isint = isint;
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

/* a runtime value (constant table, global, upvalue, argument) as a symbol */
static mt_sym mt_fromvalue(mt_ctx *c, value val) {
    if (MORPHO_ISFLOAT(val)) return mt_num(mt_constf(c, MORPHO_GETFLOATVALUE(val)));
    if (MORPHO_ISINTEGER(val)) return mt_num(mt_consti(c, MORPHO_GETINTEGERVALUE(val)));
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
static bool mt_arith(mt_ctx *c, int op, const mt_sym *l, const mt_sym *r, mt_sym *out) {
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

static bool mt_interp(mt_ctx *c, value fn, const mt_sym *args, int nargs, mt_sym *ret) {
    objectclosure *closure = NULL;
    objectfunction *func;
    if (MORPHO_ISCLOSURE(fn)) { closure = MORPHO_GETCLOSURE(fn); func = closure->func; }
    else if (MORPHO_ISFUNCTION(fn)) func = MORPHO_GETFUNCTION(fn);
    else return mt_fail(c, "integrand is not a Morpho function");
    if (func->nargs != nargs || func->varg >= 0 || func->opt.count > 0 || func->nopt > 0) return mt_fail(c, "argument count / optional arguments");
    if (++c->depth > MT_MAXDEPTH) return mt_fail(c, "call depth");
    if (!c->v->instructions) return mt_fail(c, "no program");
    mt_sym *reg = (mt_sym *) malloc(sizeof(mt_sym) * (MORPHO_MAXREGISTERS + 8));      /* INVOKE touches reg[a + 2] */
    if (!reg) return mt_fail(c, "allocation");
    for (int i = 0; i < MORPHO_MAXREGISTERS + 8; i++) reg[i] = mt_none();
    reg[0] = mt_none(); reg[0].kind = MT_OBJ; reg[0].obj = fn;
    for (int i = 0; i < nargs; i++) reg[i + 1] = args[i];
    for (int i = nargs + 1; i < func->nregs && i <= MORPHO_MAXREGISTERS; i++) reg[i] = mt_num(mt_consti(c, 0));   /* vm.c:733 */
    bool ok = false, done = false;
    const instruction *pc = c->v->instructions + func->entry;
    for (int steps = 0; !done && steps < 4096; steps++, pc++) {
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
            case OP_RETURN:
                if (a > 0) { *ret = reg[b]; ok = ret->kind == MT_NUM || ret->kind == MT_VEC; if (!ok) mt_fail(c, "function returns something that is not a number"); }
                else mt_fail(c, "function returns nil");
                done = true;
                break;
            default:
                mt_fail(c, "control flow or an instruction the translator does not handle");
                done = true;
        }
    }
    free(reg);
    c->depth--;
    return ok && !c->why;
}

/* postfix emission; shared subexpressions are re-emitted (the program has no registers) */
static bool mt_emit(mt_ctx *c, int node, mcu_expr_op *prog, int *n, int cap) {
    /* explicit stack: expressions from long sums are deep on the left */
    typedef struct { int node, stage; } frame;
    frame *st = (frame *) malloc(sizeof(frame) * (MT_MAXNODES + 8));
    if (!st) return mt_fail(c, "allocation");
    int sp = 0;
    bool ok = true;
    st[sp].node = node; st[sp++].stage = 0;
    while (sp > 0 && ok) {
        frame *f = &st[sp - 1];
        const mt_node *nd = &c->nodes[f->node];
        const bool leaf = nd->op == MCU_X_CONST || nd->op == MCU_X_POS || nd->op == MCU_X_FIELD || nd->op == MCU_X_TANGENT ||
                          nd->op == MCU_X_NORMAL || nd->op == MCU_X_GRADQ;
        const bool binary = (nd->op >= MCU_X_ADD && nd->op <= MCU_X_DIV) || nd->op == MCU_X_POW || nd->op == MCU_X_ATAN2;
        if (sp >= MT_MAXNODES) { ok = mt_fail(c, "expression too deep"); break; }
        if (!leaf && f->stage == 0) {
            f->stage = 1;
            /* POWI keeps its operand in b (a is the exponent); unary ops keep it in a */
            st[sp].node = nd->op == MCU_X_POWI ? nd->b : nd->a; st[sp++].stage = 0;
            continue;
        }
        if (binary && f->stage == 1) { f->stage = 2; st[sp].node = nd->b; st[sp++].stage = 0; continue; }
        if (*n >= cap) { ok = mt_fail(c, "integrand program too long"); break; }
        prog[*n].op = nd->op;
        prog[*n].a = ((leaf && nd->op != MCU_X_CONST) || nd->op == MCU_X_POWI) ? nd->a : 0;
        prog[*n].c = nd->op == MCU_X_CONST ? nd->c : 0.0;
        (*n)++;
        sp--;
    }
    free(st);
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
