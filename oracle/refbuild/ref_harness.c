/* ref_harness.c — OUR thin C harness around the UNMODIFIED reference library.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path may link or call this.
 *
 * It is compiled against the reference's own headers where they lie under
 * /root/reference/src and linked against oracle/_ref/libmorpho_ref.so (built
 * from the unmodified reference sources by build_ref.sh).  It lets the Python
 * tests / bench.py (cpu_baseline leg, --impl reference) drive the reference's
 * *own* builtin methods (Area.total, Area.gradient, Nematic.fieldgradient …)
 * on meshes that are handed over as flat binary arrays, and read back raw FP64
 * results (the reference's scripts only print 6 significant figures,
 * SURVEY.md F11).
 *
 * What is called in the reference:
 *   - object_newmesh                         src/geometry/mesh.c:74
 *   - mesh_newconnectivityelement            src/geometry/mesh.c:260
 *   - sparseccs_resize (+ direct cptr/rix)   src/linalg/sparse.c:339
 *   - sparsedok_insert / mesh_freezeconnectivity (the DOK path, F7)
 *   - mesh_addgrade / mesh_addconnectivityelement (edges, transposes)
 *   - object_newfield, object_newselection, selection_selectwithid
 *   - builtin_findclassfromcstring + the class's builtin methods
 *     (functional.c:6260-6280) invoked through their own C function pointers,
 *     exactly what vm.c:1303 does for OP_INVOKE.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "vm.h"      /* defines MORPHO_CORE: the real struct svm, needed for v->fp */
#include "morpho.h"
#include "classes.h"
#include "builtin.h"
#include "matrix.h"
#include "sparse.h"
#include "mesh.h"
#include "field.h"
#include "selection.h"
#include "functional.h"
#include "integrate.h"

static vm *hvm = NULL;
static program *hprog = NULL;
static char lasterr[512];

static double now_s(void) {
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

const char *refh_last_error(void) { return lasterr; }

/* Start libmorpho and make one VM whose base call frame is live (morpho_call
 * style invocation of builtins touches v->fp). */
int refh_init(int nthreads) {
    if (hvm) { morpho_setthreadnumber(nthreads); return 0; }
    morpho_initialize();
    morpho_setthreadnumber(nthreads);
    hprog = morpho_newprogram();
    compiler *c = morpho_newcompiler(hprog);
    error err; error_init(&err);
    if (!morpho_compile("var __refh = 0\n", c, false, &err)) {
        snprintf(lasterr, sizeof(lasterr), "compile failed: %s", err.msg);
        return 1;
    }
    hvm = morpho_newvm();
    if (!hvm) return 2;
    if (!morpho_run(hvm, hprog)) {
        snprintf(lasterr, sizeof(lasterr), "bootstrap run failed");
        return 3;
    }
    morpho_freecompiler(c);
    return 0;
}

void refh_set_threads(int n) { morpho_setthreadnumber(n); }
int refh_get_threads(void) { return morpho_threadnumber(); }

/* ---------------- Mesh ---------------- */

void *refh_mesh_new(int dim, int nv, const double *vert) {
    objectmesh *m = object_newmesh((unsigned) dim, (unsigned) nv, (double *) vert);
    if (m) mesh_checkconnectivity(m);
    return m;
}

int refh_mesh_set_vertices(void *mesh, const double *vert) {
    objectmesh *m = (objectmesh *) mesh;
    memcpy(m->vert->elements, vert, sizeof(double) * (size_t) m->vert->nrows * (size_t) m->vert->ncols);
    return 0;
}

double *refh_mesh_vertex_ptr(void *mesh) { return ((objectmesh *) mesh)->vert->elements; }

static int cmp_int(const void *a, const void *b) { int x = *(const int *) a, y = *(const int *) b; return (x > y) - (x < y); }

/* Fast path for large meshes: fill the CCS arrays directly.
 * The functionals only ever read s->ccs (functional.c:115-129). */
int refh_mesh_set_grade(void *mesh, int g, int nel, const int *conn, int sort) {
    objectmesh *m = (objectmesh *) mesh;
    int nvper = g + 1;
    objectsparse *s = mesh_getconnectivityelement(m, 0, (unsigned) g);
    if (!s) s = mesh_newconnectivityelement(m, 0, (unsigned) g);
    if (!s) return 1;
    if (!sparseccs_resize(&s->ccs, (int) m->vert->ncols, nel, (unsigned) ((size_t) nel * nvper), false)) return 2;
    for (int e = 0; e <= nel; e++) s->ccs.cptr[e] = e * nvper;
    memcpy(s->ccs.rix, conn, sizeof(int) * (size_t) nel * nvper);
    if (sort) for (int e = 0; e < nel; e++) qsort(s->ccs.rix + (size_t) e * nvper, nvper, sizeof(int), cmp_int);
    return 0;
}

/* Slow, fully faithful path: insert through the reference's DOK hash and let
 * it convert (and sort) to CCS itself — sparse.c:468-529.  Small meshes. */
int refh_mesh_set_grade_dok(void *mesh, int g, int nel, const int *conn) {
    objectmesh *m = (objectmesh *) mesh;
    int nvper = g + 1;
    objectsparse *s = mesh_getconnectivityelement(m, 0, (unsigned) g);
    if (!s) s = mesh_newconnectivityelement(m, 0, (unsigned) g);
    if (!s) return 1;
    for (int e = 0; e < nel; e++)
        for (int k = 0; k < nvper; k++)
            if (!sparsedok_insert(&s->dok, conn[(size_t) e * nvper + k], e, MORPHO_NIL)) return 2;
    /* dimensions as Mesh loading leaves them */
    sparsedok_setdimensions(&s->dok, (int) m->vert->ncols, nel);
    mesh_freezeconnectivity(m);
    return 0;
}

int refh_mesh_addgrade(void *mesh, int g) {
    objectmesh *m = (objectmesh *) mesh;
    objectsparse *s = mesh_addgrade(m, g);
    return s ? s->ccs.ncols : -1;
}

int refh_mesh_grade_count(void *mesh, int g) {
    objectmesh *m = (objectmesh *) mesh;
    if (g == 0) return (int) m->vert->ncols;
    objectsparse *s = mesh_getconnectivityelement(m, 0, (unsigned) g);
    if (!s) return -1;
    if (!sparse_checkformat(s, SPARSE_CCS, true, false)) return -1;
    return s->ccs.ncols;
}

/* Copy out the connectivity (row,col) of the reference as CCS: cptr[ncols+1], rix[nentries].
 * Creates raise/lower matrices on demand exactly as the functionals do
 * (mesh_addconnectivityelement, mesh.c:663-685). Returns nentries, or <0. */
int refh_mesh_get_conn(void *mesh, int row, int col, int create, int *ncols, int *cptr, int *rix, long cap) {
    objectmesh *m = (objectmesh *) mesh;
    objectsparse *s = mesh_getconnectivityelement(m, (unsigned) row, (unsigned) col);
    if (!s && create) s = mesh_addconnectivityelement(m, (unsigned) row, (unsigned) col);
    if (!s) return -1;
    if (!sparse_checkformat(s, SPARSE_CCS, true, false)) return -2;
    int ne = s->ccs.cptr[s->ccs.ncols];
    if (ncols) *ncols = s->ccs.ncols;
    if (cptr && rix) {
        if (ne > cap) return -3;
        memcpy(cptr, s->ccs.cptr, sizeof(int) * (size_t) (s->ccs.ncols + 1));
        memcpy(rix, s->ccs.rix, sizeof(int) * (size_t) ne);
    }
    return ne;
}

void refh_mesh_free(void *mesh) { if (mesh) object_free((object *) mesh); }

/* ---------------- Field (vertex based, P1) ---------------- */

/* psize==1 → scalar prototype (field.c:90-93), else a psize×1 Matrix prototype. */
void *refh_field_new(void *mesh, int psize, const double *data) {
    objectmesh *m = (objectmesh *) mesh;
    value proto = MORPHO_FLOAT(0.0);
    objectmatrix *pm = NULL;
    if (psize > 1) {
        pm = matrix_new(psize, 1, true);
        if (!pm) return NULL;
        proto = MORPHO_OBJECT(pm);
    }
    objectfield *f = object_newfield(m, proto, MORPHO_NIL, NULL);
    if (!f) return NULL;
    if (data) memcpy(f->data.elements, data, sizeof(double) * (size_t) f->data.nrows);
    return f;
}

long refh_field_size(void *field) { return (long) ((objectfield *) field)->data.nrows; }
int refh_field_set(void *field, const double *data) {
    objectfield *f = (objectfield *) field;
    memcpy(f->data.elements, data, sizeof(double) * (size_t) f->data.nrows);
    return 0;
}
void refh_field_free(void *field) { if (field) object_free((object *) field); }

/* ---------------- Selection ---------------- */

void *refh_selection_new(void *mesh) { return object_newselection((objectmesh *) mesh); }

int refh_selection_add(void *sel, int g, int n, const int *ids) {
    objectselection *s = (objectselection *) sel;
    for (int i = 0; i < n; i++) selection_selectwithid(s, g, ids[i], true);
    return 0;
}

/* ids in the reference's own iteration order (hash-slot order, functional.c:226-229) */
int refh_selection_ids(void *sel, int g, int *out, int cap) {
    objectselection *s = (objectselection *) sel;
    int n = 0;
    if ((unsigned) g >= s->ngrades) return 0;
    dictionary *d = &s->selected[g];
    if (d->count > 0) for (unsigned int k = 0; k < d->capacity; k++) {
        if (!MORPHO_ISINTEGER(d->contents[k].key)) continue;
        if (n < cap && out) out[n] = MORPHO_GETINTEGERVALUE(d->contents[k].key);
        n++;
    }
    return n;
}
void refh_selection_free(void *sel) { if (sel) object_free((object *) sel); }

/* ---------------- Functionals ---------------- */

static bool call_method(value obj, const char *name, int nargs, value *args, value *ret) {
    value method, label = builtin_internsymbolascstring((char *) name);
    if (!morpho_lookupmethod(obj, label, &method)) {
        snprintf(lasterr, sizeof(lasterr), "no method '%s'", name);
        return false;
    }
    if (!MORPHO_ISBUILTINFUNCTION(method)) {
        snprintf(lasterr, sizeof(lasterr), "method '%s' is not a builtin", name);
        return false;
    }
    objectbuiltinfunction *f = MORPHO_GETBUILTINFUNCTION(method);
    value xargs[nargs + 1];
    xargs[0] = obj;
    for (int i = 0; i < nargs; i++) xargs[i + 1] = args[i];
    hvm->fp->nopt = 0;
    error_clear(morpho_geterror(hvm));
    *ret = (f->function)(hvm, nargs, xargs);
    error *e = morpho_geterror(hvm);
    if (e->cat != ERROR_NONE) {
        snprintf(lasterr, sizeof(lasterr), "%s: %s", e->id ? e->id : "?", e->msg);
        error_clear(e);
        return false;
    }
    return true;
}

/* Create an instance of a builtin functional class.
 *   cls      - "Area", "VolumeEnclosed", "Length", "Volume", "LineCurvatureSq",
 *              "GradSq", "NormSq", "Nematic", "NematicElectric", …
 *   fields   - 0, 1 or 2 field handles passed positionally to init
 *   names/vals - numeric properties set on the instance afterwards
 *              (ksplay, ktwist, kbend, pitch, grade(int), integrandonly(bool)) —
 *              this is what the named-option branch of the init methods does
 *              (functional.c:3925-3951). */
void *refh_functional_new(const char *cls, int nfields, void **fields, int nparams, const char **names, const double *vals) {
    value klass = builtin_findclassfromcstring((char *) cls);
    if (!MORPHO_ISCLASS(klass)) { snprintf(lasterr, sizeof(lasterr), "class '%s' not found", cls); return NULL; }
    objectinstance *inst = object_newinstance(MORPHO_GETCLASS(klass));
    if (!inst) return NULL;
    value obj = MORPHO_OBJECT(inst), ret;
    value args[2];
    for (int i = 0; i < nfields && i < 2; i++) args[i] = MORPHO_OBJECT((object *) fields[i]);
    if (!call_method(obj, MORPHO_INITIALIZER_METHOD, nfields, args, &ret)) { object_free((object *) inst); return NULL; }
    for (int i = 0; i < nparams; i++) {
        value key = builtin_internsymbolascstring((char *) names[i]);
        value val = MORPHO_FLOAT(vals[i]);
        if (strcmp(names[i], "grade") == 0) val = MORPHO_INTEGER((int) vals[i]);
        if (strcmp(names[i], "integrandonly") == 0) val = (vals[i] != 0.0 ? MORPHO_TRUE : MORPHO_FALSE);
        objectinstance_setproperty(inst, key, val);
    }
    return inst;
}
/* a 1 x n Matrix as a property of the instance (EquiElement's `weight`, functional.c:2895-2907); the matrix is owned by
 * the harness for the life of the process (a few hundred doubles per golden case) */
int refh_functional_set_rowmatrix(void *f, const char *name, int n, const double *vals) {
    objectmatrix *mm = matrix_new(1, n, false);
    if (!f || !mm) { snprintf(lasterr, sizeof(lasterr), "set_rowmatrix: allocation"); return 1; }
    memcpy(mm->elements, vals, sizeof(double) * (size_t) n);
    objectinstance_setproperty((objectinstance *) f, builtin_internsymbolascstring((char *) name), MORPHO_OBJECT(mm));
    return 0;
}
void refh_functional_free(void *f) { if (f) object_free((object *) f); }

/* Evaluate method ∈ {total, integrand, gradient, fieldgradient} of a functional.
 * Arguments follow the reference call convention f.method([field,] mesh [, sel])
 * (functional.c:82-102).  Result copied to out (cap doubles). Returns number of
 * doubles written, or <0 on error (message via refh_last_error). If secs!=NULL
 * the time spent inside the reference builtin alone is stored there. */
long refh_eval(void *func, const char *method, void *mesh, void *sel, void *field, double *out, long cap, double *secs) {
    value args[3]; int nargs = 0;
    if (field) args[nargs++] = MORPHO_OBJECT((object *) field);
    if (mesh) args[nargs++] = MORPHO_OBJECT((object *) mesh);
    if (sel) args[nargs++] = MORPHO_OBJECT((object *) sel);
    value ret = MORPHO_NIL;
    double t0 = now_s();
    bool ok = call_method(MORPHO_OBJECT((object *) func), method, nargs, args, &ret);
    double t1 = now_s();
    if (secs) *secs = t1 - t0;
    if (!ok) return -1;
    long n = -2;
    if (MORPHO_ISFLOAT(ret)) {
        if (cap >= 1 && out) out[0] = MORPHO_GETFLOATVALUE(ret);
        n = 1;
    } else if (MORPHO_ISINTEGER(ret)) {
        if (cap >= 1 && out) out[0] = (double) MORPHO_GETINTEGERVALUE(ret);
        n = 1;
    } else if (MORPHO_ISMATRIX(ret)) {
        objectmatrix *mm = MORPHO_GETMATRIX(ret);
        n = (long) mm->nels;
        if (out) { if (n > cap) n = -3; else memcpy(out, mm->elements, sizeof(double) * (size_t) n); }
    } else if (MORPHO_ISFIELD(ret)) {
        objectfield *ff = MORPHO_GETFIELD(ret);
        n = (long) ff->data.nrows;
        if (out) { if (n > cap) n = -3; else memcpy(out, ff->data.elements, sizeof(double) * (size_t) n); }
    } else if (MORPHO_ISNIL(ret)) {
        snprintf(lasterr, sizeof(lasterr), "nil result (element failure without a named error)");
        n = -4;
    }
    return n;
}

/* Timing loop with no result copy: returns mean seconds per call of the
 * reference builtin; results are left to the VM's collector. */
double refh_time(void *func, const char *method, void *mesh, void *sel, void *field, int reps) {
    value args[3]; int nargs = 0;
    if (field) args[nargs++] = MORPHO_OBJECT((object *) field);
    if (mesh) args[nargs++] = MORPHO_OBJECT((object *) mesh);
    if (sel) args[nargs++] = MORPHO_OBJECT((object *) sel);
    double tot = 0.0;
    for (int r = 0; r < reps; r++) {
        value ret = MORPHO_NIL;
        double t0 = now_s();
        bool ok = call_method(MORPHO_OBJECT((object *) func), method, nargs, args, &ret);
        tot += now_s() - t0;
        if (!ok) return -1.0;
    }
    return tot / (double) (reps > 0 ? reps : 1);
}

/* functional_fdstepsize as the reference computes it (functional.c:46-58); it is a non-static
 * function of the library but has no prototype in functional.h */
double functional_fdstepsize(double x, int order);
double refh_fdstepsize(double x, int order) { return functional_fdstepsize(x, order); }


/* ---------------- Quadrature: the reference's own element integrator with a C integrand ----------------
 * integrate_integrate (integrate.c:769-800) is what LineIntegral / AreaIntegral call when no method dictionary is
 * given (functional.c:5223, 5406), with a callback that re-enters the VM for a Morpho closure.  Here the callback
 * interprets the same postfix program the CUDA library runs (include/morpho_cuda.h, mcu_expr_op): the quadrature
 * is the reference's, the integrand is ours.  Result per element = integral * element size (functional.c:5227). */
typedef struct { int op; int a; double c; } refh_expr_op;
typedef struct { int nops; const refh_expr_op *prog; } refh_prog;

static bool refh_integrand(unsigned int dim, double *lambda, double *x, unsigned int nq, value *q, void *ref, double *fout) {
    const refh_prog *pr = (const refh_prog *) ref;
    double st[64], xx[3] = {0, 0, 0};
    int sp = 0;
    for (unsigned int k = 0; k < dim && k < 3; k++) xx[k] = x[k];
    (void) lambda;
    for (int i = 0; i < pr->nops; i++) {
        const refh_expr_op *op = &pr->prog[i];
        switch (op->op) {
            case 0: st[sp++] = op->c; break;
            case 1: st[sp++] = xx[op->a]; break;
            case 2: { double v = 0.0; if ((unsigned) op->a < nq) morpho_valuetofloat(q[op->a], &v); st[sp++] = v; break; }
            case 3: sp--; st[sp - 1] = st[sp - 1] + st[sp]; break;
            case 4: sp--; st[sp - 1] = st[sp - 1] - st[sp]; break;
            case 5: sp--; st[sp - 1] = st[sp - 1] * st[sp]; break;
            case 6: sp--; st[sp - 1] = st[sp - 1] / st[sp]; break;
            case 7: st[sp - 1] = -st[sp - 1]; break;
            case 8: st[sp - 1] = pow(st[sp - 1], (double) op->a); break;
            case 9: st[sp - 1] = sqrt(st[sp - 1]); break;
            case 10: st[sp - 1] = sin(st[sp - 1]); break;
            case 11: st[sp - 1] = cos(st[sp - 1]); break;
            case 12: st[sp - 1] = exp(st[sp - 1]); break;
            case 13: st[sp - 1] = log(st[sp - 1]); break;
            case 14: st[sp - 1] = fabs(st[sp - 1]); break;
            case 15: sp--; st[sp - 1] = pow(st[sp - 1], st[sp]); break;
            case 20: sp--; st[sp - 1] = atan2(st[sp - 1], st[sp]); break;
            case 16: st[sp - 1] = tan(st[sp - 1]); break;
            case 17: st[sp - 1] = asin(st[sp - 1]); break;
            case 18: st[sp - 1] = acos(st[sp - 1]); break;
            case 19: st[sp - 1] = atan(st[sp - 1]); break;
            case 21: st[sp - 1] = sinh(st[sp - 1]); break;
            case 22: st[sp - 1] = cosh(st[sp - 1]); break;
            case 23: st[sp - 1] = tanh(st[sp - 1]); break;
            case 24: st[sp - 1] = floor(st[sp - 1]); break;
            case 25: st[sp - 1] = ceil(st[sp - 1]); break;
            case 26: st[sp - 1] = log10(st[sp - 1]); break;
            default: return false;
        }
    }
    *fout = st[sp - 1];
    return true;
}

/* out[e] = integral over element e of grade g (all elements); fields: nq flattened components, comp[i] points at the
 * nv-long array of component i at the vertices (stride given).  Returns the number of elements or -1. */
long refh_integral(void *mesh, int g, int nops, const refh_expr_op *prog, int nq, const double **comp, const int *stride, double *out) {
    objectmesh *m = (objectmesh *) mesh;
    objectsparse *s = mesh_getconnectivityelement(m, 0, (unsigned) g);
    if (!s || nq > 32) return -1;
    refh_prog pr = {nops, prog};
    int nel = s->ccs.ncols;
    for (int e = 0; e < nel; e++) {
        int nv, *vid;
        if (!sparseccs_getrowindices(&s->ccs, e, &nv, &vid) || nv != g + 1) return -1;
        double *x[4];
        value qv[4][33];
        value *q[4];
        for (int k = 0; k < nv; k++) {
            x[k] = m->vert->elements + (size_t) vid[k] * m->vert->nrows;
            for (int i = 0; i < nq; i++) qv[k][i] = MORPHO_FLOAT(comp[i][(size_t) vid[k] * stride[i]]);
            q[k] = qv[k];
        }
        double size, val;
        if (!functional_elementsize(hvm, m, g, e, nv, vid, &size)) return -1;
        /* integrate_volint builds its table of sub-element quantities from quantity[0..3] even when there are none
         * (integrate.c:707): the VM always hands it four lists, so does this harness */
        if (!integrate_integrate(refh_integrand, m->dim, (unsigned) g, x, (unsigned) nq, (nq || g == 3) ? q : NULL, &pr, &val)) {
            snprintf(lasterr, sizeof(lasterr), "integrate_integrate failed on element %d", e);
            return -1;
        }
        out[e] = val * size;
    }
    return nel;
}
