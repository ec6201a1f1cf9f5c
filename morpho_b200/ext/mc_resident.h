/* mc_resident.h — device-resident Matrix objects behind the unchanged Matrix class.
 *
 * optimize.morpho drives every line search through a handful of Matrix methods (acc, clone, inner, norm, +, -, scalar
 * multiples, setcolumn: modules/optimize.morpho:229-303, 461-486, 686-693; builtins src/linalg/matrix.c:906-1245,
 * 1494-1503).  With the functionals on the GPU, what is left of a step in the reference's own arrangement is PCIe
 * traffic: the vertex matrix up, every gradient down, for vector operations that are memory-bound anyway.
 *
 * Here a large Matrix gets a SHADOW: an mcu_matrix on the device plus two validity bits.  The accelerated methods run
 * as kernels on the shadows and leave the host array stale; the host array is brought up to date ("flush") only when
 * something that is not accelerated is about to read it — every other builtin of the VM that receives the Matrix (or
 * a Mesh whose vertex matrix it is) passes through a trampoline that does exactly that first.  Writers through
 * non-accelerated builtins invalidate the device copy, as round 1's mirror hooks did.  A Matrix that dies takes its
 * shadow with it (objecttypedefn.freefn).
 *
 * States of a shadowed matrix:  host_valid & dev_valid (coherent) | dev_valid only (device newer) | host_valid only.
 * Matrices smaller than MORPHO_CUDA_RESIDENT_MIN elements (default 32768) are never shadowed: they behave exactly as
 * without the extension (a kernel launch costs more than the BLAS call on a few hundred doubles).
 */
#ifndef MC_RESIDENT_H
#define MC_RESIDENT_H

#define MC_MAXSHADOWS 4096

typedef struct {
    objectmatrix *mat;
    mcu_matrix *d;
    bool host_valid, dev_valid;
    int colreads;                 /* single-column reads served from the device since the last flush */
} mc_shadow;

static mc_shadow shadows[MC_MAXSHADOWS];
static int nshadows = 0;
static int n_devnewer = 0;        /* shadows with !host_valid: when 0 the trampolines have nothing to do */
static long mc_resident_min = 32768;
static long mc_stat_devops = 0, mc_stat_flushes = 0;
static double mc_stat_d2h_bytes = 0.0, mc_stat_h2d_bytes = 0.0;

static void mirror_unbind_matrix(mcu_matrix *d);     /* morphocuda.c: a mesh mirror reading its vertices from d lets go */

static mc_shadow *shadow_find(const objectmatrix *mat) {
    for (int i = 0; i < nshadows; i++) if (shadows[i].mat == mat) return &shadows[i];
    return NULL;
}

static bool shadow_eligible(const objectmatrix *mat) {
    return mat && mat->nvals == 1 && mat->obj.type == OBJECT_MATRIX && (long) mat->nrows * mat->ncols >= mc_resident_min;
}

static void shadow_set_host_valid(mc_shadow *s, bool valid) {
    if (s->host_valid && !valid) n_devnewer++;
    if (!s->host_valid && valid) n_devnewer--;
    s->host_valid = valid;
}

/* new shadow; the caller says which side holds the data */
static mc_shadow *shadow_create(objectmatrix *mat, bool host_valid, bool dev_valid) {
    if (nshadows == MC_MAXSHADOWS) return NULL;
    mcu_matrix *d = mcu_matrix_create(mat->nrows, mat->ncols);
    if (!d) return NULL;
    mc_shadow *s = &shadows[nshadows++];
    s->mat = mat; s->d = d; s->host_valid = true; s->dev_valid = dev_valid; s->colreads = 0;
    shadow_set_host_valid(s, host_valid);
    return s;
}

static void shadow_drop(mc_shadow *s) {
    if (!s) return;
    mirror_unbind_matrix(s->d);
    if (!s->host_valid) n_devnewer--;
    mcu_matrix_destroy(s->d);
    *s = shadows[--nshadows];
}

/* bring the host array up to date */
static bool shadow_to_host(mc_shadow *s) {
    if (!s || s->host_valid) return true;
    if (mcu_matrix_download(s->d, s->mat->elements) != MCU_OK) return false;
    mc_stat_flushes++;
    mc_stat_d2h_bytes += 8.0 * (double) s->mat->nrows * s->mat->ncols;
    shadow_set_host_valid(s, true);
    s->colreads = 0;
    return true;
}

/* bring the device copy up to date */
static bool shadow_to_device(mc_shadow *s) {
    if (!s) return false;
    if (s->dev_valid) return true;
    if (mcu_matrix_upload(s->d, s->mat->elements) != MCU_OK) return false;
    mc_stat_h2d_bytes += 8.0 * (double) s->mat->nrows * s->mat->ncols;
    s->dev_valid = true;
    return true;
}

/* the shadow of an operand of a device operation, device copy valid; NULL when the matrix cannot be resident */
static mc_shadow *shadow_operand(objectmatrix *mat) {
    mc_shadow *s = shadow_find(mat);
    if (!s) {
        if (!shadow_eligible(mat)) return NULL;
        s = shadow_create(mat, true, false);
        if (!s) return NULL;
    }
    return shadow_to_device(s) ? s : NULL;
}

/* a fresh result Matrix that exists on the device only */
static mc_shadow *shadow_result(MatrixIdx_t nrows, MatrixIdx_t ncols, objectmatrix **out) {
    objectmatrix *new = matrix_new(nrows, ncols, false);
    if (!new) return NULL;
    mc_shadow *s = shadow_create(new, false, true);
    if (!s) { object_free((object *) new); return NULL; }
    *out = new;
    return s;
}

static void shadow_flush_value(value val);

/* a non-accelerated builtin is about to run: whatever it could read must be valid on the host */
static void shadow_flush_args(int nargs, value *args) {
    if (n_devnewer == 0) return;
    for (int i = 0; i <= nargs; i++) shadow_flush_value(args[i]);
}

static void shadow_flush_value(value val) {
    if (!MORPHO_ISOBJECT(val)) return;
    if (MORPHO_ISMATRIX(val)) { shadow_to_host(shadow_find(MORPHO_GETMATRIX(val))); return; }
    if (MORPHO_ISMESH(val)) { objectmesh *m = MORPHO_GETMESH(val); if (m->vert) shadow_to_host(shadow_find(m->vert)); return; }
    if (MORPHO_ISLIST(val)) {           /* e.g. Matrix([[a, b], [c, d]]) block constructors, functionals given lists */
        objectlist *l = MORPHO_GETLIST(val);
        if (l->val.count <= 64) for (unsigned int i = 0; i < l->val.count; i++) shadow_flush_value(l->val.data[i]);
        else for (int i = 0; i < nshadows; i++) shadow_to_host(&shadows[i]);
    }
}

/* ------------------------------------------------------------------ accelerated Matrix methods */

typedef enum { ACC_NONE = 0, ACC_ADD, ACC_SUB, ACC_MUL, ACC_DIV, ACC_ADDR, ACC_SUBR, ACC_MULR, ACC_ACC, ACC_CLONE, ACC_INNER,
               ACC_NORM, ACC_SUM, ACC_ASSIGN, ACC_SETCOLUMN, ACC_GETCOLUMN, ACC_META } mc_accel;

/* is a device operation on these operands worthwhile / necessary? */
static bool resident_pair(objectmatrix *a, objectmatrix *b) {
    mc_shadow *sa = shadow_find(a), *sb = b ? shadow_find(b) : NULL;
    if ((sa && !sa->host_valid) || (sb && !sb->host_valid)) return true;       /* the data is on the device already */
    return shadow_eligible(a) && (!b || shadow_eligible(b));
}

static value bind_result(vm *v, objectmatrix *new) {
    value out = MORPHO_OBJECT(new);
    morpho_bindobjects(v, 1, &out);
    return out;
}

/* out = sa * a + sb * b as a new Matrix */
static bool accel_axpby(vm *v, objectmatrix *a, double sa, objectmatrix *b, double sb, value *out) {
    mc_shadow *da = shadow_operand(a), *db = b ? shadow_operand(b) : NULL;
    if (!da || (b && !db)) return false;
    mcu_matrix *pa = da->d, *pb = db ? db->d : NULL;          /* shadow_result may move table entries: keep the handles */
    objectmatrix *new = NULL;
    mc_shadow *dn = shadow_result(a->nrows, a->ncols, &new);
    if (!dn) return false;
    if (mcu_matrix_axpby(dn->d, pa, sa, pb, sb) != MCU_OK) { shadow_drop(dn); object_free((object *) new); return false; }
    mc_stat_devops++;
    *out = bind_result(v, new);
    return true;
}

/* out = a * sgn + beta as a new Matrix (matrix_addscalar on a clone) */
static bool accel_xpa(vm *v, objectmatrix *a, double sgn, double beta, value *out) {
    mc_shadow *da = shadow_operand(a);
    if (!da) return false;
    mcu_matrix *pa = da->d;
    objectmatrix *new = NULL;
    mc_shadow *dn = shadow_result(a->nrows, a->ncols, &new);
    if (!dn) return false;
    if (mcu_matrix_xpa(dn->d, pa, sgn, beta) != MCU_OK) { shadow_drop(dn); object_free((object *) new); return false; }
    mc_stat_devops++;
    *out = bind_result(v, new);
    return true;
}

/* One accelerated method.  Returns true when handled (result in *out, errors raised as the reference raises them);
 * false: the caller runs the reference builtin (after flushing its arguments). */
static bool accel_call(mc_accel kind, vm *v, int nargs, value *args, value *out) {
    if (!MORPHO_ISMATRIX(MORPHO_SELF(args))) return false;
    objectmatrix *a = MORPHO_GETMATRIX(MORPHO_SELF(args));
    if (a->nvals != 1 || a->obj.type != OBJECT_MATRIX) return false;
    value arg0 = nargs > 0 ? MORPHO_GETARG(args, 0) : MORPHO_NIL;
    objectmatrix *b = (nargs > 0 && MORPHO_ISMATRIX(arg0)) ? MORPHO_GETMATRIX(arg0) : NULL;
    if (b && (b->nvals != 1 || b->obj.type != OBJECT_MATRIX)) return false;
    double x;
    *out = MORPHO_NIL;
    switch (kind) {
        case ACC_META:                                    /* count(), dimensions(): no element is read */
            return false;
        case ACC_ADD: case ACC_SUB:                       /* Matrix_add__matrix / _sub__matrix / __nil / __x (matrix.c:1125-1145) */
            if (nargs != 1) return false;
            if (MORPHO_ISNIL(arg0)) { *out = MORPHO_SELF(args); return true; }
            if (b) {
                if (!resident_pair(a, b)) return false;
                if (a->ncols != b->ncols || a->nrows != b->nrows) { morpho_runtimeerror(v, LINALG_INCOMPATIBLEMATRICES); return true; }
                return accel_axpby(v, a, 1.0, b, kind == ACC_ADD ? 1.0 : -1.0, out);
            }
            if (matrix_isamatrix(arg0)) return false;     /* another matrix type: the reference redirects to addr / subr */
            if (!resident_pair(a, NULL) || !morpho_valuetofloat(arg0, &x)) return false;
            return accel_xpa(v, a, 1.0, kind == ACC_ADD ? x : -x, out);
        case ACC_ADDR:                                    /* x + Matrix (matrix.c:1133-1136), nil + Matrix */
            if (nargs != 1) return false;
            if (MORPHO_ISNIL(arg0)) { *out = MORPHO_SELF(args); return true; }
            if (matrix_isamatrix(arg0) || !resident_pair(a, NULL) || !morpho_valuetofloat(arg0, &x)) return false;
            return accel_xpa(v, a, 1.0, x, out);
        case ACC_SUBR:                                    /* x - Matrix (matrix.c:1147-1149); -M compiles to 0 - M */
            if (nargs != 1 || !resident_pair(a, NULL) || !morpho_valuetofloat(arg0, &x)) return false;
            return accel_xpa(v, a, -1.0, x, out);
        case ACC_MUL: case ACC_MULR:                      /* Matrix * x, x * Matrix (matrix.c:1151-1160) */
            if (nargs != 1 || b || !resident_pair(a, NULL) || !morpho_valuetofloat(arg0, &x)) return false;
            return accel_axpby(v, a, x, NULL, 0.0, out);
        case ACC_DIV:                                     /* Matrix / x (matrix.c:1175-1186) */
            if (nargs != 1 || b || !resident_pair(a, NULL) || !morpho_valuetofloat(arg0, &x)) return false;
            x = 1.0 / x;
            if (isnan(x)) morpho_runtimeerror(v, VM_DVZR);
            return accel_axpby(v, a, x, NULL, 0.0, out);
        case ACC_ACC: {                                   /* a.acc(alpha, b) (matrix.c:1199-1209) */
            if (nargs != 2 || !MORPHO_ISMATRIX(MORPHO_GETARG(args, 1))) return false;
            b = MORPHO_GETMATRIX(MORPHO_GETARG(args, 1));
            if (b->nvals != 1 || b->obj.type != OBJECT_MATRIX || !resident_pair(a, b)) return false;
            if (!morpho_valuetofloat(arg0, &x)) { morpho_runtimeerror(v, LINALG_ARITHARGS); return true; }
            if (a->ncols != b->ncols || a->nrows != b->nrows) { morpho_runtimeerror(v, LINALG_INCOMPATIBLEMATRICES); return true; }
            mc_shadow *da = shadow_operand(a);
            mcu_matrix *pa = da ? da->d : NULL;
            mc_shadow *db = shadow_operand(b);
            if (!da || !db) return false;
            da = shadow_find(a);
            if (mcu_matrix_axpby(pa, pa, 1.0, db->d, x) != MCU_OK) return false;
            shadow_set_host_valid(da, false);
            mc_stat_devops++;
            return true;
        }
        case ACC_CLONE: {                                 /* matrix.c:914-918 */
            if (nargs != 0 || !resident_pair(a, NULL)) return false;
            return accel_axpby(v, a, 1.0, NULL, 0.0, out);
        }
        case ACC_ASSIGN: {                                /* a.assign(b): matrix_copy(b, a) (matrix.c:906-911) */
            if (nargs != 1 || !b || !resident_pair(a, b)) return false;
            if (a->ncols != b->ncols || a->nrows != b->nrows) { morpho_runtimeerror(v, LINALG_INCOMPATIBLEMATRICES); return true; }
            mc_shadow *db = shadow_operand(b);
            mcu_matrix *pb = db ? db->d : NULL;
            mc_shadow *da = shadow_find(a);
            if (!da && shadow_eligible(a)) da = shadow_create(a, true, false);
            if (!da || !pb) return false;
            if (mcu_matrix_copy(da->d, pb) != MCU_OK) return false;
            da->dev_valid = true;
            shadow_set_host_valid(da, false);
            mc_stat_devops++;
            return true;
        }
        case ACC_INNER: {                                 /* matrix.c:1494-1503 */
            if (nargs != 1 || !b || !resident_pair(a, b)) return false;
            if (a->ncols != b->ncols || a->nrows != b->nrows) { morpho_runtimeerror(v, LINALG_INCOMPATIBLEMATRICES); return true; }
            mc_shadow *da = shadow_operand(a);
            mcu_matrix *pa = da ? da->d : NULL;
            mc_shadow *db = (b == a) ? da : shadow_operand(b);
            if (!pa || !db) return false;
            if (mcu_matrix_inner(pa, db->d, &x) != MCU_OK) return false;
            mc_stat_devops++;
            *out = MORPHO_FLOAT(x);
            return true;
        }
        case ACC_NORM: case ACC_SUM: {                    /* Frobenius norm / sum (matrix.c:1231-1244) */
            if (nargs != 0 || !resident_pair(a, NULL)) return false;
            mc_shadow *da = shadow_operand(a);
            if (!da) return false;
            if ((kind == ACC_NORM ? mcu_matrix_norm(da->d, &x) : mcu_matrix_sum(da->d, &x)) != MCU_OK) return false;
            mc_stat_devops++;
            *out = MORPHO_FLOAT(x);
            return true;
        }
        case ACC_SETCOLUMN: {                             /* a.setcolumn(i, col) (matrix.c:1082-1087; fixgrad optimize.morpho:686-693) */
            mc_shadow *da = shadow_find(a);
            if (!da || nargs != 2 || !MORPHO_ISINTEGER(arg0) || !MORPHO_ISMATRIX(MORPHO_GETARG(args, 1))) return false;
            objectmatrix *c = MORPHO_GETMATRIX(MORPHO_GETARG(args, 1));
            int col = MORPHO_GETINTEGERVALUE(arg0);
            if (c->nvals != 1 || col < 0 || col >= a->ncols || c->nrows * c->ncols != a->nrows) return false;   /* the reference raises */
            shadow_to_host(shadow_find(c));
            /* write-through: both copies that are valid receive the column, neither is invalidated */
            if (da->dev_valid && mcu_matrix_set_column(da->d, col, c->elements) != MCU_OK) return false;
            if (da->host_valid) memcpy(a->elements + (size_t) col * a->nrows, c->elements, sizeof(double) * a->nrows);
            mc_stat_devops++;
            return true;
        }
        case ACC_GETCOLUMN: {                             /* a.column(i) (matrix.c:1068-1080) */
            mc_shadow *da = shadow_find(a);
            if (!da || da->host_valid || nargs != 1 || !MORPHO_ISINTEGER(arg0)) return false;
            int col = MORPHO_GETINTEGERVALUE(arg0);
            if (col < 0 || col >= a->ncols) return false;
            if (++da->colreads > 64) return false;        /* a loop over the columns: one flush, then host speed */
            objectmatrix *new = matrix_new(a->nrows, 1, false);
            if (!new) return false;
            if (mcu_matrix_get_column(da->d, col, new->elements) != MCU_OK) { object_free((object *) new); return false; }
            *out = bind_result(v, new);
            return true;
        }
        default: return false;
    }
}

#endif
