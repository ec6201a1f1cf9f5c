/* morphocuda.c — Morpho EXTENSION that puts libmorpho_cuda.so behind the unchanged builtin API.
 *
 * `import morphocuda` (extensions.c:134-154 dlopens <pkgroot>/lib/morphocuda.so and calls
 * morphocuda_initialize) re-points the C function pointers of the functional classes' `total`,
 * `integrand` and `gradient` builtins (objectbuiltinfunction.function, builtin.h:60-67; the class method
 * dictionaries of clss.h:19-28) at the veneers below.  Scripts, the VM, the compiler and optimize.morpho are
 * untouched: `Area().gradient(mesh)` now evaluates on the GPU and still returns a Matrix bound to the VM.
 *
 * The veneers mirror the FUNCTIONAL_* macros of functional.h:276-406: functional_validateargs, then the map,
 * then morpho_bindobjects.  Behind them a device MIRROR of every Mesh that was evaluated is kept
 * (mcu_mesh: vertex matrix + incidence of the grades used) and re-uploaded only when stale:
 *   - vertex data: the Matrix mutators (assign, setindex x3, setcolumn x2, acc, reshape — matrix.c:1590-1632) and
 *     Mesh.setvertexposition are wrapped "call the original, then mark mirrors of that matrix dirty";
 *     Mesh.setvertexmatrix swaps the pointer mesh->vert (mesh.c:1145-1160) — caught by comparing pointers
 *     at every evaluation;
 *   - connectivity: compared by Sparse object, rix pointer and element count at every evaluation; the in-place
 *     writers Sparse.setrowindices / setindex and Mesh.addgrade / removegrade / resetconnectivity invalidate;
 *   - freeing a Mesh (GC) drops its mirror (objecttypedefn.freefn wrapper, object.h:88-96).
 * MORPHO_CUDA_PARANOID=1 re-uploads the vertex matrix on every call instead of trusting the hooks.
 *
 * Anything the CUDA library does not accelerate (fields, symmetries, ragged connectivity, MCU_ERR_UNSUPPORTED)
 * is passed to the saved reference builtin — that is the reference's own code, not a fallback of ours.
 *
 * Built only where the reference headers are available (morpho_b200/ext/build_ext.sh); see INTEGRATION.md.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* MORPHO_CORE: the integrand translator reads bytecode and the VM's globals (core.h); the reference's own geometry
 * sources are compiled the same way */
#define MORPHO_CORE
#include "core.h"
#include "morpho.h"
#include "builtin.h"
#include "classes.h"
#include "matrix.h"
#include "sparse.h"
#include "mesh.h"
#include "selection.h"
#include "field.h"
#include "functional.h"
#include "integrate.h"

#include <math.h>

#include "morpho_cuda.h"
#include "mc_translate.h"
#include "mc_resident.h"

#define MC_MAXMIRRORS 64
#define MC_MAXWRAPS 1024

typedef struct {
    objectmesh *mesh;
    mcu_mesh *m;
    objectmatrix *vert;      /* mesh->vert when last mirrored */
    double *elements;        /* vert->elements when last mirrored */
    int nv, dim;
    bool vert_dirty;
    mcu_matrix *bound;       /* the device-resident Matrix (mc_resident.h) the mesh reads its vertices from, or NULL */
    struct { objectsparse *s; int *rix; int nel; bool ok; } conn[4];
} mc_mirror;

static mc_mirror mirrors[MC_MAXMIRRORS];
static int nmirrors = 0;
static bool mc_ready = false, mc_paranoid = false, mc_verbose = false, mc_dry = false;
static long mc_stat_evals = 0, mc_stat_vert_uploads = 0, mc_stat_conn_uploads = 0, mc_stat_fallbacks = 0;

static void mc_log(const char *what, const char *detail) {
    if (mc_verbose) fprintf(stderr, "[morphocuda] %s %s\n", what, detail ? detail : "");
}

/* ------------------------------------------------------------------ mirrors */

static void fields_drop_for(mcu_mesh *m);
static void mirror_drop(int i) {
    if (mirrors[i].m) { fields_drop_for(mirrors[i].m); mcu_mesh_destroy(mirrors[i].m); }
    mirrors[i] = mirrors[nmirrors - 1];
    nmirrors--;
}

static mc_mirror *mirror_find(objectmesh *mesh) {
    for (int i = 0; i < nmirrors; i++) if (mirrors[i].mesh == mesh) return &mirrors[i];
    return NULL;
}

static void field_mark_range(const double *elements);
static void mirror_mark_matrix(objectmatrix *a) {
    for (int i = 0; i < nmirrors; i++)
        if (mirrors[i].vert == a || mirrors[i].elements == a->elements) mirrors[i].vert_dirty = true;
    field_mark_range(a->elements);
}

/* a device-resident Matrix is going away (mc_resident.h shadow_drop): mirrors reading from it return to their own buffer */
static void mirror_unbind_matrix(mcu_matrix *d) {
    for (int i = 0; i < nmirrors; i++)
        if (mirrors[i].bound == d) { mcu_mesh_bind_vertices(mirrors[i].m, NULL); mirrors[i].bound = NULL; mirrors[i].vert_dirty = true; }
}

static void mirror_invalidate_connectivity(objectmesh *mesh) {
    for (int i = 0; i < nmirrors; i++)
        if (!mesh || mirrors[i].mesh == mesh) for (int g = 0; g < 4; g++) mirrors[i].conn[g].ok = false;
}

/* Mirror of `mesh`, coherent for grade g (0 = vertices only).  NULL → caller uses the reference builtin. */
static mc_mirror *mirror_get(objectmesh *mesh, grade g) {
    if (!mesh->vert || mesh->vert->nvals != 1) return NULL;
    int dim = mesh->vert->nrows, nv = mesh->vert->ncols;
    if (dim != (int) mesh->dim || (dim != 2 && dim != 3)) return NULL;
    mc_mirror *mr = mirror_find(mesh);
    if (mr && (mr->nv != nv || mr->dim != dim)) { mirror_drop((int) (mr - mirrors)); mr = NULL; }
    if (!mr) {
        if (nmirrors == MC_MAXMIRRORS) mirror_drop(0);
        mr = &mirrors[nmirrors];
        memset(mr, 0, sizeof(*mr));
        mr->m = mcu_mesh_create(dim, nv);
        if (!mr->m) return NULL;
        mr->mesh = mesh; mr->nv = nv; mr->dim = dim; mr->vert_dirty = true;
        nmirrors++;
    }
    if (mr->vert != mesh->vert || mr->elements != mesh->vert->elements) {   /* Mesh.setvertexmatrix swapped the matrix */
        mr->vert = mesh->vert; mr->elements = mesh->vert->elements; mr->vert_dirty = true;
    }
    /* a vertex matrix large enough to live on the device (mc_resident.h) is READ IN PLACE: the mirror binds to the
     * Matrix's shadow, so `x.acc(-s, force)` of a line search is seen without any copy and Mesh.setvertexmatrix is a
     * pointer swap here as it is in the reference (mesh.c:1145-1160) */
    mc_shadow *vs = shadow_find(mesh->vert);
    if (!vs && shadow_eligible(mesh->vert) && !mc_paranoid) vs = shadow_create(mesh->vert, true, false);
    if (vs && !mc_paranoid) {
        if (!vs->dev_valid) mc_stat_vert_uploads++;
        if (!shadow_to_device(vs)) return NULL;
        if (mr->bound != vs->d) {
            if (mcu_mesh_bind_vertices(mr->m, vs->d) != MCU_OK) return NULL;
            mr->bound = vs->d;
        }
        mr->vert_dirty = false;
    } else {
        if (vs && !shadow_to_host(vs)) return NULL;
        if (mr->bound) { mcu_mesh_bind_vertices(mr->m, NULL); mr->bound = NULL; mr->vert_dirty = true; }
        if (mr->vert_dirty || mc_paranoid) {
            if (mcu_mesh_set_vertices(mr->m, mesh->vert->elements) != MCU_OK) return NULL;
            mr->vert_dirty = false;
            mc_stat_vert_uploads++;
        }
    }
    if (g > 0) {
        objectsparse *s = mesh_getconnectivityelement(mesh, 0, g);
        if (!s) return mr;                                   /* caller raises FnctlELNtFnd */
        if (!sparse_checkformat(s, SPARSE_CCS, true, false)) return NULL;
        int nel = s->ccs.ncols;
        if (!mr->conn[g].ok || mr->conn[g].s != s || mr->conn[g].rix != s->ccs.rix || mr->conn[g].nel != nel) {
            /* the device mirror stores rix with an implicit cptr[e] = e (g+1): check the simplicial layout */
            if (s->ccs.nentries != nel * (g + 1)) return NULL;
            for (int e = 0; e <= nel; e++) if (s->ccs.cptr[e] != e * (g + 1)) return NULL;
            for (int k = 0; k < s->ccs.nentries; k++) if (s->ccs.rix[k] < 0 || s->ccs.rix[k] >= nv) return NULL;
            if (mcu_mesh_set_connectivity(mr->m, g, nel, s->ccs.rix) != MCU_OK) return NULL;
            mr->conn[g].s = s; mr->conn[g].rix = s->ccs.rix; mr->conn[g].nel = nel; mr->conn[g].ok = true;
            mc_stat_conn_uploads++;
        }
    }
    return mr;
}

/* ------------------------------------------------------------------ field mirrors */

/* Vertex (P1) fields only: one entry per vertex, psize doubles each, no finite-element space (field.c:419-434).
 * The device copy is refreshed only when the VM wrote to the Field: its own writers (Field.setindex / assign / acc,
 * field.c:1136-1144) are wrapped, and writes through the Matrix objects that ALIAS its store — the pool matrices
 * Field_getindex hands out (field.c:309-327) and the store itself from __linearize (field.c:1128-1132) — are caught
 * by address range in the Matrix writer hooks (field_mark_range).  A Field that dies takes its mirrors with it.
 * MORPHO_CUDA_PARANOID=1 uploads on every evaluation instead. */
#define MC_MAXFIELDS 32
typedef struct { objectfield *fld; mcu_mesh *m; mcu_field *f; int psize, nv; long used; double *elements; bool dirty; } mc_fieldmirror;
static mc_fieldmirror fields[MC_MAXFIELDS];
static int nfields = 0;
static long mc_stat_field_uploads = 0;
static long mc_call_gen = 0;    /* one evaluation = one generation: entries fetched by the current call are never evicted */

static void fields_drop_for(mcu_mesh *m) {
    for (int i = 0; i < nfields;) {
        if (!m || fields[i].m == m) { mcu_field_destroy(fields[i].f); fields[i] = fields[--nfields]; }
        else i++;
    }
}

static bool field_is_vertex_field(objectfield *fld, objectmesh *mesh) {
    if (!fld || fld->mesh != mesh || !MORPHO_ISNIL(fld->fnspc) || fld->ngrades < 1 || fld->dof[0] != 1) return false;
    for (unsigned int g = 1; g < fld->ngrades; g++) if (fld->dof[g] != 0) return false;
    if (fld->psize < 1 || fld->psize > 9) return false;
    return fld->data.nels == (MatrixCount_t) mesh->vert->ncols * fld->psize;
}

static void field_mark(objectfield *fld) {
    for (int i = 0; i < nfields; i++) if (fields[i].fld == fld) fields[i].dirty = true;
}

/* a Matrix at `elements` was written: it may alias the store of a mirrored Field */
static void field_mark_range(const double *elements) {
    for (int i = 0; i < nfields; i++)
        if (elements >= fields[i].elements && elements < fields[i].elements + (size_t) fields[i].nv * fields[i].psize) fields[i].dirty = true;
}

static void fields_drop_field(objectfield *fld) {
    for (int i = 0; i < nfields;) {
        if (fields[i].fld == fld) { mcu_field_destroy(fields[i].f); fields[i] = fields[--nfields]; }
        else i++;
    }
}

static mcu_field *field_get(objectfield *fld, mc_mirror *mr) {
    mc_fieldmirror *fm = NULL;
    for (int i = 0; i < nfields; i++)
        if (fields[i].fld == fld && fields[i].m == mr->m && fields[i].psize == (int) fld->psize && fields[i].nv == mr->nv) fm = &fields[i];
    if (!fm) {
        if (nfields == MC_MAXFIELDS) {
            /* evict the least recently used entry that the CURRENT evaluation has not fetched (an evaluation holds up
             * to 16 mcu_field pointers while it fetches the rest) */
            int victim = -1;
            for (int i = 0; i < nfields; i++)
                if (fields[i].used != mc_call_gen && (victim < 0 || fields[i].used < fields[victim].used)) victim = i;
            if (victim < 0) return NULL;                       /* table full of this call's own fields: reference path */
            mcu_field_destroy(fields[victim].f);
            fields[victim] = fields[--nfields];
        }
        fm = &fields[nfields];
        fm->f = mcu_field_create(mr->m, (int) fld->psize);
        if (!fm->f) return NULL;
        fm->fld = fld; fm->m = mr->m; fm->psize = (int) fld->psize; fm->nv = mr->nv;
        fm->elements = NULL; fm->dirty = true;
        nfields++;
    }
    fm->used = mc_call_gen;
    /* the store handed out by __linearize may have become a device-resident Matrix that was written there */
    mc_shadow *fs = nshadows > 0 ? shadow_find(&fld->data) : NULL;
    if (fs && !fs->host_valid) { if (!shadow_to_host(fs)) return NULL; fm->dirty = true; }
    if (fm->dirty || fm->elements != fld->data.elements || mc_paranoid) {
        if (mcu_field_set(fm->f, fld->data.elements) != MCU_OK) return NULL;
        fm->elements = fld->data.elements; fm->dirty = false;
        mc_stat_field_uploads++;
    }
    return fm->f;
}

/* ------------------------------------------------------------------ self-check mode */

/* MORPHO_CUDA_VERIFY=1: every evaluation the device handled is ALSO run through the saved reference builtin on the same
 * inputs; the two results are compared (scalars: relative difference; matrices and fields: max |difference| over max
 * |reference|), the worst difference per Class.method is reported at exit, and the REFERENCE's result is what the script
 * receives — so a long optimisation follows the stock trajectory exactly while each of its thousands of evaluations
 * is checked.  (Optimisers amplify last-digit differences of finite-difference gradients: examples/tactoid reaches
 * different refined meshes even between `-w0` and `-w4` of the stock reference, so end results cannot be compared.) */
#define MC_MAXVERIFY 64
static bool mc_verify = false;
static struct { char name[48]; long n, mismatched; double worst; } mc_verify_tab[MC_MAXVERIFY];
static int mc_nverify = 0;

static double mc_verify_arrays(const double *a, const double *b, size_t n) {
    double num = 0.0, den = 0.0;
    for (size_t i = 0; i < n; i++) { double d = fabs(a[i] - b[i]); if (d > num || d != d) num = d; if (fabs(b[i]) > den) den = fabs(b[i]); }
    if (num != num) return INFINITY;
    return den > 0.0 ? num / den : (num == 0.0 ? 0.0 : INFINITY);
}

static void mc_verify_record(const char *cls, const char *meth, value ours, value ref) {
    double err = INFINITY;
    double x, y;
    if (MORPHO_ISMATRIX(ours) && MORPHO_ISMATRIX(ref)) {
        shadow_flush_value(ours); shadow_flush_value(ref);
        objectmatrix *a = MORPHO_GETMATRIX(ours), *b = MORPHO_GETMATRIX(ref);
        if (a->nrows == b->nrows && a->ncols == b->ncols) err = mc_verify_arrays(a->elements, b->elements, (size_t) a->nrows * a->ncols);
    } else if (MORPHO_ISFIELD(ours) && MORPHO_ISFIELD(ref)) {
        objectfield *a = MORPHO_GETFIELD(ours), *b = MORPHO_GETFIELD(ref);
        if (a->data.nrows == b->data.nrows) err = mc_verify_arrays(a->data.elements, b->data.elements, (size_t) a->data.nrows);
    } else if (morpho_valuetofloat(ours, &x) && morpho_valuetofloat(ref, &y)) {
        err = fabs(x - y) / (fabs(y) > 1e-300 ? fabs(y) : 1e-300);
        if (x == y) err = 0.0;
    } else if (MORPHO_ISNIL(ours) && MORPHO_ISNIL(ref)) err = 0.0;
    char name[48];
    snprintf(name, sizeof(name), "%s.%s", cls, meth);
    int k = 0;
    while (k < mc_nverify && strcmp(mc_verify_tab[k].name, name)) k++;
    if (k == mc_nverify && mc_nverify < MC_MAXVERIFY) { memset(&mc_verify_tab[k], 0, sizeof(mc_verify_tab[k])); strcpy(mc_verify_tab[k].name, name); mc_nverify++; }
    if (k < mc_nverify) {
        mc_verify_tab[k].n++;
        if (!(err <= mc_verify_tab[k].worst)) mc_verify_tab[k].worst = err;
        if (!(err < INFINITY)) mc_verify_tab[k].mismatched++;
    }
}

static value mc_verified(vm *v, const char *cls, const char *meth, value ours, builtinfunction orig, int nargs, value *args) {
    int handle = MORPHO_ISOBJECT(ours) ? morpho_retainobjects(v, 1, &ours) : -1;
    value ref = orig(v, nargs, args);
    mc_verify_record(cls, meth, ours, ref);
    if (handle >= 0) morpho_releaseobjects(v, handle);
    return ref;
}

static void mc_verify_report(void) {
    for (int k = 0; k < mc_nverify; k++)
        fprintf(stderr, "[morphocuda verify] %s evaluations %ld worst %.3e shape-or-type mismatches %ld\n", mc_verify_tab[k].name,
                mc_verify_tab[k].n, mc_verify_tab[k].worst, mc_verify_tab[k].mismatched);
}

/* ------------------------------------------------------------------ the veneers */

/* ids of a Selection for grade g: the integer keys of sel->selected[g] (functional.c:226-229) */
static int selection_ids(objectselection *sel, grade g, int **out) {
    *out = NULL;
    if ((unsigned) g >= sel->ngrades) return 0;
    dictionary *d = &sel->selected[g];
    int n = 0;
    int *ids = malloc(sizeof(int) * (d->count > 0 ? d->count : 1));
    if (!ids) return -1;
    if (d->count > 0) for (unsigned int k = 0; k < d->capacity; k++)
        if (MORPHO_ISINTEGER(d->contents[k].key)) ids[n++] = MORPHO_GETINTEGERVALUE(d->contents[k].key);
    *out = ids;
    return n;
}

static bool mesh_has_symmetries(objectmesh *mesh, grade g) {
    return mesh_getconnectivityelement(mesh, 0, 0) || (g > 0 && mesh_getconnectivityelement(mesh, g, g));
}

/* A Matrix-valued result (gradient / integrand).  Large results stay on the device: the objectmatrix the builtin
 * returns is shadowed by the mcu_matrix the kernels wrote (mc_resident.h) and its host array is filled only if something
 * that is not accelerated reads it.  Small ones are copied to the host at once, as the reference would have them. */
static int eval_matrix_result(mc_mirror *mr, const mcu_functional *f, mcu_selection *sel, int mode, MatrixIdx_t nr, MatrixIdx_t nc, objectmatrix **out) {
    objectmatrix *new = matrix_new(nr, nc, false);
    *out = new;
    if (!new) return -1;
    if ((long) nr * nc == 0) return MCU_OK;
    if (shadow_eligible(new) && !mc_paranoid) {
        mc_shadow *s = shadow_create(new, false, true);
        if (s) {
            int rc = mcu_eval_matrix(mr->m, f, sel, mode, s->d);
            if (rc == MCU_OK) rc = mcu_mesh_status(mr->m);          /* degenerate element / VolEnclZero raised by the kernels */
            if (rc != MCU_OK) shadow_drop(shadow_find(new));
            else mc_stat_devops++;
            return rc;
        }
    }
    return mcu_eval(mr->m, f, sel, mode, new->elements);
}

/* One evaluation; returns true when handled (out/err set), false → call the reference builtin */
static bool mc_eval(vm *v, int nargs, value *args, int functional, grade g, int mode, value *out) {
    functional_mapinfo info;
    *out = MORPHO_NIL;
    if (!mc_ready) return false;
    if (!functional_validateargs(v, nargs, args, &info)) return true;    /* FnctlIntMsh raised, result nil */
    if (info.field || mesh_has_symmetries(info.mesh, g)) return false;
    if (info.sel && info.sel->mode != SELECT_SOME) return false;
    mc_mirror *mr = mirror_get(info.mesh, g);
    if (!mr) return false;
    if (g > 0 && !mesh_getconnectivityelement(info.mesh, 0, g)) {
        morpho_runtimeerror(v, FUNC_ELNTFND, g);
        return true;
    }
    mcu_functional f;
    memset(&f, 0, sizeof(f));
    f.functional = functional; f.grade = g;
    mcu_selection *sel = NULL;
    if (info.sel) {
        int *ids, n = selection_ids(info.sel, g, &ids);
        if (n < 0) return false;
        sel = mcu_selection_create(mr->m, g, n, ids);
        free(ids);
        if (!sel) return false;
    }
    int rc = MCU_OK;
    objectmatrix *new = NULL;
    double total = 0.0;
    switch (mode) {
        case MCU_TOTAL:
            rc = mcu_eval(mr->m, &f, sel, MCU_TOTAL, &total);
            break;
        case MCU_INTEGRAND: {
            long n = mcu_result_size(mr->m, &f, MCU_INTEGRAND);
            rc = eval_matrix_result(mr, &f, sel, MCU_INTEGRAND, 1, (MatrixIdx_t) n, &new);
            if (rc == -1) morpho_runtimeerror(v, ERROR_ALLOCATIONFAILED);
            break;
        }
        case MCU_GRADIENT:
            rc = eval_matrix_result(mr, &f, sel, MCU_GRADIENT, mr->dim, mr->nv, &new);
            if (rc == -1) morpho_runtimeerror(v, ERROR_ALLOCATIONFAILED);
            break;
        default:
            rc = MCU_ERR_UNSUPPORTED;
    }
    if (sel) mcu_selection_destroy(sel);
    mc_stat_evals++;
    if (rc == MCU_OK) {
        if (mode == MCU_TOTAL) *out = MORPHO_FLOAT(total);
        else { *out = MORPHO_OBJECT(new); morpho_bindobjects(v, 1, out); }
        return true;
    }
    if (new) object_free((object *) new);
    if (rc == -1) return true;                                /* allocation error already raised */
    if (rc == MCU_ERR_DEGENERATE) return true;                /* reference: map aborted, method result nil (functional.c:392,412-415) */
    if (rc == MCU_ERR_VOLENCLZERO) { morpho_runtimeerror(v, VOLUMEENCLOSED_ZERO); return true; }
    if (rc == MCU_ERR_ELNOTFOUND) { morpho_runtimeerror(v, FUNC_ELNTFND, g); return true; }
    mc_stat_evals--;
    mc_log("not accelerated:", mcu_last_error());
    return false;
}

/* Functionals that read their parameters from instance properties (the reference's *_prepareref functions:
 * curvature 2940-2961, gradsq 3623-3642, nematic 3757-3791, nematicelectric 4008-4033). */
static bool prop(objectinstance *self, const char *name, value *out) {
    objectstring key = MORPHO_STATICSTRING((char *) name);
    return objectinstance_getproperty(self, MORPHO_OBJECT(&key), out);
}

static bool mc_eval_props(vm *v, int nargs, value *args, int functional, int mode, value *out) {
    functional_mapinfo info;
    mc_call_gen++;
    *out = MORPHO_NIL;
    if (!mc_ready || !MORPHO_ISINSTANCE(MORPHO_SELF(args))) return false;
    if (!functional_validateargs(v, nargs, args, &info)) return true;
    objectinstance *self = MORPHO_GETINSTANCE(MORPHO_SELF(args));
    if (info.sel && info.sel->mode != SELECT_SOME) return false;
    mcu_functional f;
    memset(&f, 0, sizeof(f));
    f.functional = functional;
    f.ksplay = f.ktwist = f.kbend = 1.0;
    value val;
    objectfield *own[2] = {NULL, NULL};
    grade g = 0;
    if (functional == MCU_F_LINECURVATURESQ) {
        g = 0;
        if (prop(self, CURVATURE_INTEGRANDONLY_PROPERTY, &val)) f.integrandonly = MORPHO_ISTRUE(val);
        if (mesh_has_symmetries(info.mesh, 1)) return false;
    } else {
        /* the field(s) the functional was constructed with */
        if (!prop(self, FUNCTIONAL_FIELD_PROPERTY, &val)) return false;
        if (functional == MCU_F_NEMATICELECTRIC) {
            if (!MORPHO_ISLIST(val)) return false;
            value a = MORPHO_NIL, b = MORPHO_NIL;
            list_getelement(MORPHO_GETLIST(val), 0, &a);
            list_getelement(MORPHO_GETLIST(val), 1, &b);
            if (!MORPHO_ISFIELD(a) || !MORPHO_ISFIELD(b)) return false;      /* a Matrix as the electric field: reference path */
            own[0] = MORPHO_GETFIELD(a); own[1] = MORPHO_GETFIELD(b);
        } else {
            if (!MORPHO_ISFIELD(val)) return false;
            own[0] = MORPHO_GETFIELD(val);
        }
        if (functional == MCU_F_NORMSQ) g = 0;
        else {
            g = mesh_maxgrade(info.mesh);
            if (prop(self, FUNCTIONAL_GRADE_PROPERTY, &val) && MORPHO_ISINTEGER(val) && MORPHO_GETINTEGERVALUE(val) > 0) g = MORPHO_GETINTEGERVALUE(val);
            if (g != 2 && g != 3) return false;
        }
        if (functional == MCU_F_NEMATIC) {
            if (prop(self, NEMATIC_KSPLAY_PROPERTY, &val) && MORPHO_ISNUMBER(val)) morpho_valuetofloat(val, &f.ksplay);
            if (prop(self, NEMATIC_KTWIST_PROPERTY, &val) && MORPHO_ISNUMBER(val)) morpho_valuetofloat(val, &f.ktwist);
            if (prop(self, NEMATIC_KBEND_PROPERTY, &val) && MORPHO_ISNUMBER(val)) morpho_valuetofloat(val, &f.kbend);
            if (prop(self, NEMATIC_PITCH_PROPERTY, &val) && MORPHO_ISNUMBER(val)) { morpho_valuetofloat(val, &f.pitch); f.haspitch = 1; }
        }
        for (int i = 0; i < 2; i++) if (own[i] && !field_is_vertex_field(own[i], info.mesh)) return false;
        if (mesh_has_symmetries(info.mesh, g)) return false;
    }
    f.grade = g;
    mc_mirror *mr = mirror_get(info.mesh, functional == MCU_F_LINECURVATURESQ ? 1 : g);
    if (!mr) return false;
    if (functional == MCU_F_LINECURVATURESQ && !mesh_getconnectivityelement(info.mesh, 0, 1)) return false;
    if (g > 0 && !mesh_getconnectivityelement(info.mesh, 0, g)) { morpho_runtimeerror(v, FUNC_ELNTFND, g); return true; }
    for (int i = 0; i < 2; i++) if (own[i]) { f.field[i] = field_get(own[i], mr); if (!f.field[i]) return false; }
    objectfield *wrtfield = NULL;
    if (mode == MCU_FIELDGRADIENT) {
        /* the Field argument is the one differentiated; it has to be one of the functional's own (functional.c:1454-1460) */
        if (!info.field) return false;
        if (info.field == own[0]) f.wrt = 0;
        else if (info.field == own[1]) f.wrt = 1;
        else return false;
        wrtfield = info.field;
    }
    mcu_selection *sel = NULL;
    if (info.sel) {
        int *ids, n = selection_ids(info.sel, g, &ids);
        if (n < 0) return false;
        sel = mcu_selection_create(mr->m, g, n, ids);
        free(ids);
        if (!sel) return false;
    }
    int rc = MCU_OK;
    objectmatrix *new = NULL;
    objectfield *newf = NULL;
    double total = 0.0;
    long n = mcu_result_size(mr->m, &f, mode);
    if (n < 0) rc = MCU_ERR_UNSUPPORTED;
    else switch (mode) {
        case MCU_TOTAL: rc = mcu_eval(mr->m, &f, sel, MCU_TOTAL, &total); break;
        case MCU_INTEGRAND:
            rc = eval_matrix_result(mr, &f, sel, MCU_INTEGRAND, 1, (MatrixIdx_t) n, &new);
            if (rc == -1) morpho_runtimeerror(v, ERROR_ALLOCATIONFAILED);
            break;
        case MCU_GRADIENT:
            rc = eval_matrix_result(mr, &f, sel, MCU_GRADIENT, mr->dim, mr->nv, &new);
            if (rc == -1) morpho_runtimeerror(v, ERROR_ALLOCATIONFAILED);
            break;
        case MCU_FIELDGRADIENT:
            newf = object_newfield(info.mesh, wrtfield->prototype, wrtfield->fnspc, wrtfield->dof);   /* functional.c:1449 */
            if (!newf || (long) newf->data.nels != n) { if (newf) object_free((object *) newf); newf = NULL; morpho_runtimeerror(v, ERROR_ALLOCATIONFAILED); rc = -1; break; }
            rc = mcu_eval(mr->m, &f, sel, MCU_FIELDGRADIENT, newf->data.elements);
            break;
        default: rc = MCU_ERR_UNSUPPORTED;
    }
    if (sel) mcu_selection_destroy(sel);
    if (rc == MCU_OK) {
        mc_stat_evals++;
        if (mode == MCU_TOTAL) *out = MORPHO_FLOAT(total);
        else { *out = MORPHO_OBJECT(new ? (object *) new : (object *) newf); morpho_bindobjects(v, 1, out); }
        return true;
    }
    if (new) object_free((object *) new);
    if (newf) object_free((object *) newf);
    if (rc == -1) return true;
    if (rc == MCU_ERR_ELNOTFOUND) { morpho_runtimeerror(v, FUNC_ELNTFND, g); return true; }
    mc_log("not accelerated:", mcu_last_error());
    return false;
}

#define MC_PVENEER(cls, FID, meth, MODE) \
    static builtinfunction orig_##cls##_##meth = NULL; \
    static value mc_##cls##_##meth(vm *v, int nargs, value *args) { \
        value out; \
        if (mc_eval_props(v, nargs, args, FID, MODE, &out)) return mc_verify ? mc_verified(v, #cls, #meth, out, orig_##cls##_##meth, nargs, args) : out; \
        mc_stat_fallbacks++; \
        return orig_##cls##_##meth(v, nargs, args); \
    }
#define MC_FIELDFUNCTIONAL(cls, FID) \
    MC_PVENEER(cls, FID, total, MCU_TOTAL) \
    MC_PVENEER(cls, FID, integrand, MCU_INTEGRAND) \
    MC_PVENEER(cls, FID, gradient, MCU_GRADIENT) \
    MC_PVENEER(cls, FID, fieldgradient, MCU_FIELDGRADIENT)

MC_FIELDFUNCTIONAL(GradSq, MCU_F_GRADSQ)
MC_FIELDFUNCTIONAL(NormSq, MCU_F_NORMSQ)
MC_FIELDFUNCTIONAL(Nematic, MCU_F_NEMATIC)
MC_FIELDFUNCTIONAL(NematicElectric, MCU_F_NEMATICELECTRIC)
MC_PVENEER(LineCurvatureSq, MCU_F_LINECURVATURESQ, total, MCU_TOTAL)
MC_PVENEER(LineCurvatureSq, MCU_F_LINECURVATURESQ, integrand, MCU_INTEGRAND)
MC_PVENEER(LineCurvatureSq, MCU_F_LINECURVATURESQ, gradient, MCU_GRADIENT)

/* EquiElement total / integrand / gradient (functional.c:2775-2922; the regulariser of examples/tactoid/tactoid.morpho:16).
 * Properties: grade (-1: highest grade of the mesh) and weight (a 1 x nel Matrix or nil).  Failures of the map
 * (a vertex whose elements have zero mean size raises EquiElArgs in the reference) go to the reference builtin. */
static bool mc_eval_equielement(vm *v, int nargs, value *args, int mode, value *out) {
    functional_mapinfo info;
    *out = MORPHO_NIL;
    if (!mc_ready || !MORPHO_ISINSTANCE(MORPHO_SELF(args))) return false;
    if (!functional_validateargs(v, nargs, args, &info)) return true;
    objectinstance *self = MORPHO_GETINSTANCE(MORPHO_SELF(args));
    if (info.field || (info.sel && info.sel->mode != SELECT_SOME)) return false;
    value val = MORPHO_NIL;
    if (!prop(self, FUNCTIONAL_GRADE_PROPERTY, &val) || !MORPHO_ISINTEGER(val)) return false;       /* equielement_prepareref fails: EquiElArgs */
    int eg = MORPHO_GETINTEGERVALUE(val), maxg = (int) mesh_maxgrade(info.mesh);
    if (eg < 0 || eg > maxg) eg = maxg;
    if (eg < 1 || eg > 3 || mesh_has_symmetries(info.mesh, eg)) return false;
    /* the reference leaves conn(grade, 0) behind in the mesh (mesh_addconnectivityelement, functional.c:2791): so do we */
    if (!mesh_addconnectivityelement(info.mesh, eg, 0) || !mesh_getconnectivityelement(info.mesh, 0, eg)) return false;
    mcu_functional f;
    memset(&f, 0, sizeof(f));
    f.functional = MCU_F_EQUIELEMENT; f.grade = eg;
    if (prop(self, EQUIELEMENT_WEIGHT_PROPERTY, &val) && MORPHO_ISMATRIX(val)) {
        objectmatrix *w = MORPHO_GETMATRIX(val);
        if (w->nvals != 1 || w->nrows != 1) return false;          /* matrix_getelement(weight, 0, el): a row of weights */
        if (!shadow_to_host(shadow_find(w))) return false;
        f.host_weight = w->elements; f.nweight = (int) w->ncols;
    }
    mc_mirror *mr = mirror_get(info.mesh, eg);
    if (!mr) return false;
    mcu_selection *sel = NULL;
    if (info.sel) {
        int *ids, n = selection_ids(info.sel, 0, &ids);
        if (n < 0) return false;
        sel = mcu_selection_create(mr->m, 0, n, ids);
        free(ids);
        if (!sel) return false;
    }
    int rc = MCU_OK;
    objectmatrix *new = NULL;
    double total = 0.0;
    switch (mode) {
        case MCU_TOTAL: rc = mcu_eval(mr->m, &f, sel, MCU_TOTAL, &total); break;
        case MCU_INTEGRAND: rc = eval_matrix_result(mr, &f, sel, MCU_INTEGRAND, 1, mr->nv, &new); break;
        case MCU_GRADIENT: rc = eval_matrix_result(mr, &f, sel, MCU_GRADIENT, mr->dim, mr->nv, &new); break;
        default: rc = MCU_ERR_UNSUPPORTED;
    }
    if (sel) mcu_selection_destroy(sel);
    if (rc == MCU_OK) {
        mc_stat_evals++;
        if (mode == MCU_TOTAL) *out = MORPHO_FLOAT(total);
        else { *out = MORPHO_OBJECT(new); morpho_bindobjects(v, 1, out); }
        return true;
    }
    if (new) object_free((object *) new);
    if (rc == -1) { morpho_runtimeerror(v, ERROR_ALLOCATIONFAILED); return true; }
    mc_log("EquiElement not accelerated:", mcu_last_error());
    return false;
}
#define MC_EVENEER(meth, MODE) \
    static builtinfunction orig_EquiElement_##meth = NULL; \
    static value mc_EquiElement_##meth(vm *v, int nargs, value *args) { \
        value out; \
        if (mc_eval_equielement(v, nargs, args, MODE, &out)) return mc_verify ? mc_verified(v, "EquiElement", #meth, out, orig_EquiElement_##meth, nargs, args) : out; \
        mc_stat_fallbacks++; \
        return orig_EquiElement_##meth(v, nargs, args); \
    }
MC_EVENEER(total, MCU_TOTAL)
MC_EVENEER(integrand, MCU_INTEGRAND)
MC_EVENEER(gradient, MCU_GRADIENT)

/* LineIntegral / AreaIntegral / VolumeIntegral (functional.c:5186-5243, 5371-5426, 5471-5546).  The integrand closure is
 * translated into a postfix program (mc_translate.h); what cannot be translated — and reference meshes, fields on
 * finite-element spaces — goes to the reference builtin.  A `method` dictionary selects the rule-table integrator. */
#define MC_MAXPROG 256
static long mc_stat_translated = 0;
static bool mc_eval_integral(vm *v, int nargs, value *args, grade g, int mode, value *out) {
    functional_mapinfo info;
    mc_call_gen++;
    *out = MORPHO_NIL;
    if (!mc_ready || !MORPHO_ISINSTANCE(MORPHO_SELF(args))) return false;
    if (!functional_validateargs(v, nargs, args, &info)) return true;
    objectinstance *self = MORPHO_GETINSTANCE(MORPHO_SELF(args));
    if ((info.field != NULL) != (mode == MCU_FIELDGRADIENT)) return false;
    if ((info.sel && info.sel->mode != SELECT_SOME) || mesh_has_symmetries(info.mesh, g)) return false;
    value fn = MORPHO_NIL, val = MORPHO_NIL;
    int wrt = -1;
    if (!prop(self, SCALARPOTENTIAL_FUNCTION_PROPERTY, &fn)) return false;
    /* a `method` dictionary selects the rule-table integrator (integrator_configurewithdictionary, integrate.c:2706-2751):
     * "rule" (String), "degree" (Int), "adapt" (Bool); an entry of the wrong type is the reference's error to raise */
    mcu_quad_method qm, *method = NULL;
    if (prop(self, INTEGRAL_METHOD_PROPERTY, &val) && !MORPHO_ISNIL(val)) {
        if (!MORPHO_ISDICTIONARY(val)) return false;
        objectdictionary *md = MORPHO_GETDICTIONARY(val);
        memset(&qm, 0, sizeof(qm));
        qm.adapt = 1; qm.degree = -1;
        objectstring krule = MORPHO_STATICSTRING(INTEGRATE_RULELABEL), kdeg = MORPHO_STATICSTRING(INTEGRATE_DEGREELABEL), kadapt = MORPHO_STATICSTRING(INTEGRATE_ADAPTLABEL);
        value e;
        if (dictionary_get(&md->dict, MORPHO_OBJECT(&krule), &e)) {
            if (!MORPHO_ISSTRING(e) || strlen(MORPHO_GETCSTRING(e)) >= sizeof(qm.rule)) return false;
            strcpy(qm.rule, MORPHO_GETCSTRING(e));
        }
        if (dictionary_get(&md->dict, MORPHO_OBJECT(&kdeg), &e)) { if (!MORPHO_ISINTEGER(e)) return false; qm.degree = MORPHO_GETINTEGERVALUE(e); }
        if (dictionary_get(&md->dict, MORPHO_OBJECT(&kadapt), &e)) { if (!MORPHO_ISBOOL(e)) return false; qm.adapt = MORPHO_GETBOOLVALUE(e) ? 1 : 0; }
        method = &qm;
    }
    if (prop(self, LINEARELASTICITY_REFERENCE_PROPERTY, &val) && !MORPHO_ISNIL(val)) return false;
    if (prop(self, LINEARELASTICITY_WTBYREF_PROPERTY, &val) && MORPHO_ISTRUE(val)) return false;
    objectfield *flds[16];
    int nf = 0, rows[16], cols[16];
    if (prop(self, FUNCTIONAL_FIELD_PROPERTY, &val) && MORPHO_ISLIST(val)) {
        objectlist *l = MORPHO_GETLIST(val);
        if (l->val.count > 16) return false;
        for (unsigned int i = 0; i < l->val.count; i++) {
            if (!MORPHO_ISFIELD(l->val.data[i])) return false;
            objectfield *f = MORPHO_GETFIELD(l->val.data[i]);
            if (!field_is_vertex_field(f, info.mesh)) return false;
            if (MORPHO_ISMATRIX(f->prototype)) {
                objectmatrix *pm = MORPHO_GETMATRIX(f->prototype);
                if (pm->nvals != 1 || (unsigned) (pm->nrows * pm->ncols) != f->psize) return false;
                rows[nf] = (int) pm->nrows; cols[nf] = (int) pm->ncols;
            } else if ((MORPHO_ISNIL(f->prototype) || MORPHO_ISNUMBER(f->prototype)) && f->psize == 1) { rows[nf] = 0; cols[nf] = 0; }
            else return false;
            if (f == info.field) wrt = nf;
            flds[nf++] = f;
        }
    }
    if (mode == MCU_FIELDGRADIENT && wrt < 0) return false;      /* not one of the functional's fields: reference path */
    mc_mirror *mr = mirror_get(info.mesh, g);
    if (!mr) return false;
    if (!mesh_getconnectivityelement(info.mesh, 0, g)) { morpho_runtimeerror(v, FUNC_ELNTFND, g); return true; }
    mcu_expr_op prog[MC_MAXPROG];
    int nops = 0;
    const char *why = "";
    value fvals[16];
    for (int i = 0; i < nf; i++) fvals[i] = MORPHO_OBJECT(flds[i]);
    if (!mc_translate_integrand(v, fn, mr->dim, nf, rows, cols, fvals, prog, MC_MAXPROG, &nops, &why)) {
        mc_log("integrand not translated:", why);
        return false;
    }
    mc_stat_translated++;
    mcu_field *df[16];
    for (int i = 0; i < nf; i++) { df[i] = field_get(flds[i], mr); if (!df[i]) return false; }
    mcu_selection *sel = NULL;
    if (info.sel) {
        int *ids, n = selection_ids(info.sel, g, &ids);
        if (n < 0) return false;
        sel = mcu_selection_create(mr->m, g, n, ids);
        free(ids);
        if (!sel) return false;
    }
    int rc;
    double total = 0.0;
    objectmatrix *new = NULL;
    objectfield *newf = NULL;
    if (mode == MCU_TOTAL) rc = mcu_integral_method(mr->m, g, nops, prog, nf, df, sel, MCU_TOTAL, method, &total);
    else if (mode == MCU_INTEGRAND) {
        objectsparse *s = mesh_getconnectivityelement(info.mesh, 0, g);
        new = matrix_new(1, (MatrixIdx_t) s->ccs.ncols, false);
        if (!new) { morpho_runtimeerror(v, ERROR_ALLOCATIONFAILED); rc = -1; }
        else if (s->ccs.ncols > 0) rc = mcu_integral_method(mr->m, g, nops, prog, nf, df, sel, MCU_INTEGRAND, method, new->elements);
        else rc = MCU_OK;
    } else if (mode == MCU_GRADIENT) {
        new = matrix_new(mr->dim, mr->nv, false);
        if (!new) { morpho_runtimeerror(v, ERROR_ALLOCATIONFAILED); rc = -1; }
        else rc = mcu_integral_gradient_method(mr->m, g, nops, prog, nf, df, sel, -1, method, new->elements);
    } else {
        objectfield *w = flds[wrt];
        newf = object_newfield(info.mesh, w->prototype, w->fnspc, w->dof);           /* functional.c:1449 */
        if (!newf || newf->data.nels != w->data.nels) { if (newf) object_free((object *) newf); newf = NULL; morpho_runtimeerror(v, ERROR_ALLOCATIONFAILED); rc = -1; }
        else rc = mcu_integral_gradient_method(mr->m, g, nops, prog, nf, df, sel, wrt, method, newf->data.elements);
    }
    if (sel) mcu_selection_destroy(sel);
    if (rc == MCU_OK) {
        mc_stat_evals++;
        if (mode == MCU_TOTAL) *out = MORPHO_FLOAT(total);
        else { *out = MORPHO_OBJECT(new ? (object *) new : (object *) newf); morpho_bindobjects(v, 1, out); }
        return true;
    }
    if (new) object_free((object *) new);
    if (newf) object_free((object *) newf);
    if (rc == -1) return true;
    mc_log("integral not accelerated:", mcu_last_error());
    return false;
}

/* ScalarPotential total / integrand / gradient (functional.c:2255-2372): the potential (and, if given, its gradient
 * function) is translated like an integrand; without a gradient function the device takes the same central
 * differences as functional_mapnumericalgradient. */
static bool mc_eval_scalarpotential(vm *v, int nargs, value *args, int mode, value *out) {
    functional_mapinfo info;
    *out = MORPHO_NIL;
    if (!mc_ready || !MORPHO_ISINSTANCE(MORPHO_SELF(args))) return false;
    if (!functional_validateargs(v, nargs, args, &info)) return true;
    objectinstance *self = MORPHO_GETINSTANCE(MORPHO_SELF(args));
    if (info.field || (info.sel && info.sel->mode != SELECT_SOME) || mesh_has_symmetries(info.mesh, 0)) return false;
    value fn = MORPHO_NIL, gfn = MORPHO_NIL;
    if (!prop(self, SCALARPOTENTIAL_FUNCTION_PROPERTY, &fn)) return false;
    bool hasgrad = prop(self, SCALARPOTENTIAL_GRADFUNCTION_PROPERTY, &gfn) && !MORPHO_ISNIL(gfn);
    mc_mirror *mr = mirror_get(info.mesh, 0);
    if (!mr) return false;
    mcu_expr_op prog[MC_MAXPROG];
    int nops[3] = {0, 0, 0}, nprog = 0;
    const char *why = "";
    if (!mc_translate_pointfn(v, (mode == MCU_GRADIENT && hasgrad) ? gfn : fn, mr->dim, prog, MC_MAXPROG, nops, &nprog, &why) ||
        nprog != ((mode == MCU_GRADIENT && hasgrad) ? mr->dim : 1)) {
        mc_log("potential not translated:", why);
        return false;
    }
    mc_stat_translated++;
    mcu_selection *sel = NULL;
    if (info.sel) {
        int *ids, n = selection_ids(info.sel, 0, &ids);
        if (n < 0) return false;
        sel = mcu_selection_create(mr->m, 0, n, ids);
        free(ids);
        if (!sel) return false;
    }
    int rc;
    double total = 0.0;
    objectmatrix *new = NULL;
    if (mode == MCU_TOTAL) rc = mcu_pointwise(mr->m, nprog, nops, prog, sel, MCU_TOTAL, &total);
    else {
        new = (mode == MCU_INTEGRAND) ? matrix_new(1, mr->nv, false) : matrix_new(mr->dim, mr->nv, false);
        if (!new) { morpho_runtimeerror(v, ERROR_ALLOCATIONFAILED); rc = -1; }
        else rc = mcu_pointwise(mr->m, nprog, nops, prog, sel, mode, new->elements);
    }
    if (sel) mcu_selection_destroy(sel);
    if (rc == MCU_OK) {
        mc_stat_evals++;
        if (mode == MCU_TOTAL) *out = MORPHO_FLOAT(total);
        else { *out = MORPHO_OBJECT(new); morpho_bindobjects(v, 1, out); }
        return true;
    }
    if (new) object_free((object *) new);
    if (rc == -1) return true;
    mc_log("potential not accelerated:", mcu_last_error());
    return false;
}
#define MC_SVENEER(meth, MODE) \
    static builtinfunction orig_ScalarPotential_##meth = NULL; \
    static value mc_ScalarPotential_##meth(vm *v, int nargs, value *args) { \
        value out; \
        if (mc_eval_scalarpotential(v, nargs, args, MODE, &out)) return mc_verify ? mc_verified(v, "ScalarPotential", #meth, out, orig_ScalarPotential_##meth, nargs, args) : out; \
        mc_stat_fallbacks++; \
        return orig_ScalarPotential_##meth(v, nargs, args); \
    }
MC_SVENEER(total, MCU_TOTAL)
MC_SVENEER(integrand, MCU_INTEGRAND)
MC_SVENEER(gradient, MCU_GRADIENT)

#define MC_IVENEER(cls, GRADE, meth, MODE) \
    static builtinfunction orig_##cls##_##meth = NULL; \
    static value mc_##cls##_##meth(vm *v, int nargs, value *args) { \
        value out; \
        if (mc_eval_integral(v, nargs, args, GRADE, MODE, &out)) return mc_verify ? mc_verified(v, #cls, #meth, out, orig_##cls##_##meth, nargs, args) : out; \
        mc_stat_fallbacks++; \
        return orig_##cls##_##meth(v, nargs, args); \
    }
MC_IVENEER(LineIntegral, MESH_GRADE_LINE, total, MCU_TOTAL)
MC_IVENEER(LineIntegral, MESH_GRADE_LINE, integrand, MCU_INTEGRAND)
MC_IVENEER(LineIntegral, MESH_GRADE_LINE, gradient, MCU_GRADIENT)
MC_IVENEER(LineIntegral, MESH_GRADE_LINE, fieldgradient, MCU_FIELDGRADIENT)
MC_IVENEER(AreaIntegral, MESH_GRADE_AREA, total, MCU_TOTAL)
MC_IVENEER(AreaIntegral, MESH_GRADE_AREA, integrand, MCU_INTEGRAND)
MC_IVENEER(AreaIntegral, MESH_GRADE_AREA, gradient, MCU_GRADIENT)
MC_IVENEER(AreaIntegral, MESH_GRADE_AREA, fieldgradient, MCU_FIELDGRADIENT)
MC_IVENEER(VolumeIntegral, MESH_GRADE_VOLUME, total, MCU_TOTAL)
MC_IVENEER(VolumeIntegral, MESH_GRADE_VOLUME, integrand, MCU_INTEGRAND)
MC_IVENEER(VolumeIntegral, MESH_GRADE_VOLUME, gradient, MCU_GRADIENT)
MC_IVENEER(VolumeIntegral, MESH_GRADE_VOLUME, fieldgradient, MCU_FIELDGRADIENT)

#define MC_VENEER(cls, FID, GRADE, meth, MODE) \
    static builtinfunction orig_##cls##_##meth = NULL; \
    static value mc_##cls##_##meth(vm *v, int nargs, value *args) { \
        value out; \
        if (mc_eval(v, nargs, args, FID, GRADE, MODE, &out)) return mc_verify ? mc_verified(v, #cls, #meth, out, orig_##cls##_##meth, nargs, args) : out; \
        mc_stat_fallbacks++; \
        return orig_##cls##_##meth(v, nargs, args); \
    }
#define MC_FUNCTIONAL(cls, FID, GRADE) \
    MC_VENEER(cls, FID, GRADE, total, MCU_TOTAL) \
    MC_VENEER(cls, FID, GRADE, integrand, MCU_INTEGRAND) \
    MC_VENEER(cls, FID, GRADE, gradient, MCU_GRADIENT)

MC_FUNCTIONAL(Length, MCU_F_LENGTH, MESH_GRADE_LINE)
MC_FUNCTIONAL(Area, MCU_F_AREA, MESH_GRADE_AREA)
MC_FUNCTIONAL(VolumeEnclosed, MCU_F_VOLUMEENCLOSED, MESH_GRADE_AREA)
MC_FUNCTIONAL(Volume, MCU_F_VOLUME, MESH_GRADE_VOLUME)
MC_VENEER(AreaEnclosed, MCU_F_AREAENCLOSED, MESH_GRADE_LINE, total, MCU_TOTAL)
MC_VENEER(AreaEnclosed, MCU_F_AREAENCLOSED, MESH_GRADE_LINE, integrand, MCU_INTEGRAND)

#include "mc_pairwise.h"

/* ------------------------------------------------------------------ hooks on the VM's builtins */

/* Every builtin the extension does not replace outright passes through a numbered trampoline:
 *   accel   an accelerated Matrix method (mc_resident.h) is tried first;
 *   flush   before the reference builtin runs, Matrix arguments (and the vertex matrix of Mesh arguments) whose
 *           newest copy is on the device are brought to the host;
 *   post    what the builtin wrote: the receiver Matrix (its device copy and the mirrors reading it are stale), a
 *           mesh's vertices, connectivity. */
typedef enum { WRAP_NONE, WRAP_MATRIX_SELF, WRAP_MATRIX_RESHAPE, WRAP_MESH_VERTS, WRAP_MESH_CONN, WRAP_CONN_ALL, WRAP_FIELD_SELF } mc_wrapkind;
static struct { builtinfunction orig; mc_wrapkind kind; mc_accel accel; bool flush; } wraps[MC_MAXWRAPS];
static int nwraps = 0;

static value wrap_call(int slot, vm *v, int nargs, value *args) {
    value self = MORPHO_SELF(args), r;
    pairwise_scan(v);
    if (wraps[slot].accel != ACC_NONE && accel_call(wraps[slot].accel, v, nargs, args, &r)) return r;
    if (wraps[slot].flush) shadow_flush_args(nargs, args);
    /* resolve the target BEFORE the call (the method may free or replace things), mark AFTER it */
    mc_wrapkind kind = wraps[slot].kind;
    objectmatrix *mat = ((kind == WRAP_MATRIX_SELF || kind == WRAP_MATRIX_RESHAPE) && MORPHO_ISMATRIX(self)) ? MORPHO_GETMATRIX(self) : NULL;
    objectmesh *mesh = ((kind == WRAP_MESH_VERTS || kind == WRAP_MESH_CONN) && MORPHO_ISMESH(self)) ? MORPHO_GETMESH(self) : NULL;
    objectfield *fld = (kind == WRAP_FIELD_SELF && MORPHO_ISFIELD(self)) ? MORPHO_GETFIELD(self) : NULL;
    r = wraps[slot].orig(v, nargs, args);
    switch (kind) {
        case WRAP_NONE: break;
        case WRAP_FIELD_SELF: if (fld) field_mark(fld); break;
        case WRAP_MATRIX_SELF: case WRAP_MATRIX_RESHAPE:
            if (mat) {
                mirror_mark_matrix(mat);
                mc_shadow *sh = shadow_find(mat);                 /* written on the host (flushed above): device copy stale */
                if (sh && kind == WRAP_MATRIX_RESHAPE) shadow_drop(sh);
                else if (sh) sh->dev_valid = false;
            }
            break;
        case WRAP_MESH_VERTS: {
            mc_mirror *mr = mesh ? mirror_find(mesh) : NULL;
            if (mr) mr->vert_dirty = true;
            if (mesh && mesh->vert && wraps[slot].flush) { mc_shadow *sh = shadow_find(mesh->vert); if (sh) sh->dev_valid = false; }
            break;
        }
        case WRAP_MESH_CONN: mirror_invalidate_connectivity(mesh); break;
        case WRAP_CONN_ALL: mirror_invalidate_connectivity(NULL); break;
    }
    return r;
}
/* C has no closures: a pool of numbered trampolines, one per wrapped builtin */
#define TR(n) static value tr_##n(vm *v, int nargs, value *args) { return wrap_call(0x##n, v, nargs, args); }
#define TR16(p) TR(p##0) TR(p##1) TR(p##2) TR(p##3) TR(p##4) TR(p##5) TR(p##6) TR(p##7) TR(p##8) TR(p##9) TR(p##a) TR(p##b) TR(p##c) TR(p##d) TR(p##e) TR(p##f)
#define TR256(p) TR16(p##0) TR16(p##1) TR16(p##2) TR16(p##3) TR16(p##4) TR16(p##5) TR16(p##6) TR16(p##7) TR16(p##8) TR16(p##9) TR16(p##a) TR16(p##b) TR16(p##c) TR16(p##d) TR16(p##e) TR16(p##f)
TR256(0) TR256(1) TR256(2) TR256(3)
#define TRN(n) tr_##n,
#define TRN16(p) TRN(p##0) TRN(p##1) TRN(p##2) TRN(p##3) TRN(p##4) TRN(p##5) TRN(p##6) TRN(p##7) TRN(p##8) TRN(p##9) TRN(p##a) TRN(p##b) TRN(p##c) TRN(p##d) TRN(p##e) TRN(p##f)
#define TRN256(p) TRN16(p##0) TRN16(p##1) TRN16(p##2) TRN16(p##3) TRN16(p##4) TRN16(p##5) TRN16(p##6) TRN16(p##7) TRN16(p##8) TRN16(p##9) TRN16(p##a) TRN16(p##b) TRN16(p##c) TRN16(p##d) TRN16(p##e) TRN16(p##f)
static builtinfunction trampolines[MC_MAXWRAPS] = { TRN256(0) TRN256(1) TRN256(2) TRN256(3) };

static int trampoline_index(builtinfunction fn) {
    for (int k = 0; k < MC_MAXWRAPS; k++) if (fn == trampolines[k]) return k;
    return -1;
}

/* every builtin implementation behind class.method (a plain builtin or the members of a metafunction) */
static int method_impls(const char *cls, const char *method, objectbuiltinfunction **out, int cap) {
    value klass = builtin_findclassfromcstring((char *) cls);
    if (!MORPHO_ISCLASS(klass)) return 0;
    objectstring key = MORPHO_STATICSTRING((char *) method);
    value m;
    if (!dictionary_get(&MORPHO_GETCLASS(klass)->methods, MORPHO_OBJECT(&key), &m)) return 0;
    int n = 0;
    if (MORPHO_ISBUILTINFUNCTION(m)) {
        if (n < cap) out[n++] = MORPHO_GETBUILTINFUNCTION(m);
    } else if (MORPHO_ISMETAFUNCTION(m)) {
        objectmetafunction *mf = MORPHO_GETMETAFUNCTION(m);
        for (unsigned int i = 0; i < mf->fns.count; i++)
            if (MORPHO_ISBUILTINFUNCTION(mf->fns.data[i]) && n < cap) out[n++] = MORPHO_GETBUILTINFUNCTION(mf->fns.data[i]);
    }
    return n;
}

/* wrap one builtin implementation (or refine the wrap it already has) */
static int hook_impl(objectbuiltinfunction *impl, mc_wrapkind kind, mc_accel accel, bool flush) {
    int k = trampoline_index(impl->function);
    if (k >= 0) {                                             /* wrapped already (generic pass, or registered under two names) */
        if (kind != WRAP_NONE) wraps[k].kind = kind;
        if (accel != ACC_NONE) wraps[k].accel = accel;
        wraps[k].flush = flush;
        return 0;
    }
    if (nwraps == MC_MAXWRAPS) { fprintf(stderr, "[morphocuda] out of trampolines\n"); return 0; }
    wraps[nwraps].orig = impl->function; wraps[nwraps].kind = kind; wraps[nwraps].accel = accel; wraps[nwraps].flush = flush;
    impl->function = trampolines[nwraps++];
    return 1;
}

static int hook_method(const char *cls, const char *method, mc_wrapkind kind, mc_accel accel, bool flush) {
    objectbuiltinfunction *impl[16];
    int n = method_impls(cls, method, impl, 16), done = 0;
    for (int i = 0; i < n; i++) done += hook_impl(impl[i], kind, accel, flush);
    return done;
}
#define hook_writers(cls, method, kind) hook_method(cls, method, kind, ACC_NONE, true)

/* The generic pass: every method of every builtin class except plain containers and scalars, and every builtin
 * function, flushes its Matrix / Mesh arguments first.  (Lists, tuples, dictionaries, strings and ranges never read
 * the elements of a Matrix they hold; printing goes through Matrix.print, which is a method and is wrapped.) */
static bool class_is_plain(objectclass *klass) {
    static const char *skip[] = {"Object", "List", "Tuple", "Dictionary", "String", "Range", "Int", "Float", "Bool", "Nil", "Function",
                                 "Closure", "Invocation", "Error", "Class", "Metafunction", "BuiltinFunction", "Upvalue", "Instance",
                                 "DictionaryEntry", "Complex", "File", "Folder", "System", "JSON", NULL};
    if (!MORPHO_ISSTRING(klass->name)) return false;
    for (int i = 0; skip[i]; i++) if (strcmp(MORPHO_GETCSTRING(klass->name), skip[i]) == 0) return true;
    return false;
}

static int hook_value(value m) {
    int done = 0;
    if (MORPHO_ISBUILTINFUNCTION(m)) done += hook_impl(MORPHO_GETBUILTINFUNCTION(m), WRAP_NONE, ACC_NONE, true);
    else if (MORPHO_ISMETAFUNCTION(m)) {
        objectmetafunction *mf = MORPHO_GETMETAFUNCTION(m);
        for (unsigned int i = 0; i < mf->fns.count; i++)
            if (MORPHO_ISBUILTINFUNCTION(mf->fns.data[i])) done += hook_impl(MORPHO_GETBUILTINFUNCTION(mf->fns.data[i]), WRAP_NONE, ACC_NONE, true);
    }
    return done;
}

static value mc_stats(vm *v, int nargs, value *args);
static value mc_translate_fn(vm *v, int nargs, value *args);
/* while an extension initialises, builtin_getclasstable() / builtin_getfunctiontable() are the extension's own (empty)
 * tables (extensions.c:140-156); the VM's are these globals of builtin.c:27-30 */
extern dictionary builtin_functiontable, builtin_classtable;

static int hook_everything(void) {
    int done = 0;
    dictionary *classes = &builtin_classtable, *fns = &builtin_functiontable;
    for (unsigned int i = 0; i < classes->capacity; i++) {
        if (MORPHO_ISNIL(classes->contents[i].key) || !MORPHO_ISCLASS(classes->contents[i].val)) continue;
        objectclass *klass = MORPHO_GETCLASS(classes->contents[i].val);
        if (class_is_plain(klass)) continue;
        for (unsigned int k = 0; k < klass->methods.capacity; k++)
            if (!MORPHO_ISNIL(klass->methods.contents[k].key)) done += hook_value(klass->methods.contents[k].val);
    }
    for (unsigned int i = 0; i < fns->capacity; i++) {
        if (MORPHO_ISNIL(fns->contents[i].key)) continue;
        value fv = fns->contents[i].val;
        /* our own helpers are registered in this table too */
        if (MORPHO_ISBUILTINFUNCTION(fv) && (MORPHO_GETBUILTINFUNCTION(fv)->function == mc_stats || MORPHO_GETBUILTINFUNCTION(fv)->function == mc_translate_fn)) continue;
        done += hook_value(fv);
    }
    return done;
}

static int patch_method(const char *cls, const char *method, builtinfunction ours, builtinfunction *orig) {
    objectbuiltinfunction *impl[4];
    int n = method_impls(cls, method, impl, 4);
    if (n != 1) { fprintf(stderr, "[morphocuda] %s.%s not found as a single builtin; left alone\n", cls, method); return 0; }
    if (impl[0]->function == ours) return 1;
    *orig = impl[0]->function;
    impl[0]->function = ours;
    return 1;
}

/* Mesh objects that die take their mirror with them */
static objectfreefn orig_meshfree = NULL;
static void mc_meshfree(object *obj) {
    mc_mirror *mr = mirror_find((objectmesh *) obj);
    if (mr) mirror_drop((int) (mr - mirrors));
    if (orig_meshfree) orig_meshfree(obj);
}

/* Field objects that die take their device mirrors with them (a later Field may be allocated at the same address) */
static objectfreefn orig_fieldfree = NULL;
static void mc_fieldfree(object *obj) {
    if (nfields > 0) fields_drop_field((objectfield *) obj);
    if (orig_fieldfree) orig_fieldfree(obj);
}

/* Matrix objects that die take their device shadow with them (objectmatrixdefn.freefn is NULL in the reference) */
static objectfreefn orig_matrixfree = NULL;
static void mc_matrixfree(object *obj) {
    if (nshadows > 0) { mc_shadow *s = shadow_find((objectmatrix *) obj); if (s) shadow_drop(s); }
    if (orig_matrixfree) orig_matrixfree(obj);
}

/* ------------------------------------------------------------------ script-visible helpers */

/* cudastats() → List [evaluations on the GPU, vertex uploads, connectivity uploads, calls left to the reference, field uploads,
 * integrand closures translated, Matrix operations run on device-resident matrices, matrices copied back to the host for a
 * reader that is not accelerated, bytes host→device and device→host moved for resident matrices] */
static value mc_stats(vm *v, int nargs, value *args) {
    value vals[10] = {MORPHO_INTEGER((int) mc_stat_evals), MORPHO_INTEGER((int) mc_stat_vert_uploads),
                      MORPHO_INTEGER((int) mc_stat_conn_uploads), MORPHO_INTEGER((int) mc_stat_fallbacks),
                      MORPHO_INTEGER((int) mc_stat_field_uploads), MORPHO_INTEGER((int) mc_stat_translated),
                      MORPHO_INTEGER((int) mc_stat_devops), MORPHO_INTEGER((int) mc_stat_flushes),
                      MORPHO_FLOAT(mc_stat_h2d_bytes), MORPHO_FLOAT(mc_stat_d2h_bytes)};
    objectlist *l = object_newlist(10, vals);
    if (!l) return MORPHO_NIL;
    return morpho_wrapandbind(v, (object *) l);
}

/* cudatranslate(f, dim, shapes): the postfix program the integrand f(x, q...) translates to, as a List of [op, a, c]
 * triples, or a String saying why it cannot be translated.  shapes: one entry per quantity, 0 for a scalar, n for an
 * n x 1 column vector, [r, c] for a matrix.  Works without a GPU (diagnostics, tests). */
static value mc_translate_fn(vm *v, int nargs, value *args) {
    if (nargs != 3 || !MORPHO_ISINTEGER(MORPHO_GETARG(args, 1)) || !MORPHO_ISLIST(MORPHO_GETARG(args, 2))) {
        morpho_runtimeerror(v, VM_INVALIDARGS, 3, nargs);
        return MORPHO_NIL;
    }
    objectlist *sh = MORPHO_GETLIST(MORPHO_GETARG(args, 2));
    int rows[16], cols[16], nf = 0;
    for (unsigned int i = 0; i < sh->val.count && nf < 16; i++, nf++) {
        value e = sh->val.data[i];
        if (MORPHO_ISINTEGER(e)) { rows[nf] = MORPHO_GETINTEGERVALUE(e); cols[nf] = rows[nf] ? 1 : 0; }
        else if (MORPHO_ISLIST(e) && MORPHO_GETLIST(e)->val.count == 2 && MORPHO_ISINTEGER(MORPHO_GETLIST(e)->val.data[0]) && MORPHO_ISINTEGER(MORPHO_GETLIST(e)->val.data[1])) {
            rows[nf] = MORPHO_GETINTEGERVALUE(MORPHO_GETLIST(e)->val.data[0]); cols[nf] = MORPHO_GETINTEGERVALUE(MORPHO_GETLIST(e)->val.data[1]);
        } else { morpho_runtimeerror(v, VM_INVALIDARGS, 3, nargs); return MORPHO_NIL; }
    }
    mcu_expr_op prog[MC_MAXPROG];
    int nops = 0;
    const char *why = "";
    if (!mc_translate_integrand(v, MORPHO_GETARG(args, 0), MORPHO_GETINTEGERVALUE(MORPHO_GETARG(args, 1)), nf, rows, cols, NULL, prog, MC_MAXPROG, &nops, &why)) {
        value s = object_stringfromcstring(why, strlen(why));
        return morpho_wrapandbind(v, MORPHO_GETOBJECT(s));
    }
    objectlist *out = object_newlist(0, NULL);
    if (!out) return MORPHO_NIL;
    value keep[MC_MAXPROG + 1];
    int nkeep = 0;
    for (int i = 0; i < nops; i++) {
        value t[3] = {MORPHO_INTEGER(prog[i].op), MORPHO_INTEGER(prog[i].a), MORPHO_FLOAT(prog[i].c)};
        objectlist *row = object_newlist(3, t);
        if (!row) break;
        list_append(out, MORPHO_OBJECT(row));
        keep[nkeep++] = MORPHO_OBJECT(row);
    }
    keep[nkeep++] = MORPHO_OBJECT(out);
    morpho_bindobjects(v, nkeep, keep);
    return MORPHO_OBJECT(out);
}

/* ------------------------------------------------------------------ extension entry points */


/* ------------------------------------------------------------------ connectivity construction
 * Mesh.addgrade(g) (mesh_addgrade, mesh.c:419-489: an n-tuple hash, ~4 us per candidate) and Selection(mesh, boundary=true)
 * (selection_selectboundary, selection.c:211-232, which first builds the raise matrix with mesh_addlowermatrix,
 * mesh.c:621-660) on the device for meshes large enough to matter (MORPHO_CUDA_CONNECT_MIN elements of the source grade,
 * default 20000): mcu_connect.cu produces the same element ids in the same order, bit for bit.  The new incidence is
 * handed to the reference as a compressed-column Sparse (sparseccs_resize, sparse.c:339-359) — the format
 * mesh_freezeconnectivity converts every connectivity element into anyway. */
static long mc_connect_min = 20000;
bool selection_selectelement(objectselection *sel, grade g, elementid id);      /* selection.c:92, 129; mesh.c:96, 273 */
bool selection_addgradelower(objectselection *sel, grade g);
bool mesh_setconnectivityelement(objectmesh *mesh, unsigned int row, unsigned int col, objectsparse *el);
void mesh_link(objectmesh *mesh, object *obj);
static value mc_boundaryoption;       /* the interned symbol "boundary" (selection.c:568) */

static builtinfunction orig_Mesh_addgrade = NULL;
static value mc_Mesh_addgrade(vm *v, int nargs, value *args) {
    objectmesh *mesh = MORPHO_GETMESH(MORPHO_SELF(args));
    while (mc_ready && nargs == 1 && MORPHO_ISINTEGER(MORPHO_GETARG(args, 0))) {
        int g = MORPHO_GETINTEGERVALUE(MORPHO_GETARG(args, 0));
        if (g < 1 || g > 2 || mesh_getconnectivityelement(mesh, 0, (unsigned) g)) break;        /* nothing to build, or not ours */
        int maxg = (int) mesh_maxgrade(mesh), h;
        objectsparse *el = NULL;
        for (h = g + 1; h <= maxg && h <= 3; h++) if ((el = mesh_getconnectivityelement(mesh, 0, (unsigned) h))) break;
        if (!el || !sparse_checkformat(el, SPARSE_CCS, true, false) || el->ccs.ncols < mc_connect_min) { mc_log("addgrade on the reference:", "source grade absent or below MORPHO_CUDA_CONNECT_MIN"); break; }
        mc_mirror *mr = mirror_get(mesh, (grade) h);
        if (!mr || !mr->conn[h].ok) { mc_log("addgrade on the reference:", "no mirror of the source grade"); break; }
        for (int k = g; k < h; k++) { mcu_mesh_drop_connectivity(mr->m, k); mr->conn[k].ok = false; }   /* grades the Mesh no longer has */
        int nel = 0;
        if (mcu_mesh_addgrade(mr->m, g, &nel) != MCU_OK) { mc_log("addgrade not accelerated:", mcu_last_error()); break; }
        objectsparse *new = object_newsparse(NULL, NULL);
        if (!new) break;
        if (!sparseccs_resize(&new->ccs, mr->nv, nel, (unsigned) ((size_t) nel * (g + 1)), false) ||
            (nel > 0 && mcu_mesh_get_connectivity(mr->m, g, new->ccs.rix) != MCU_OK)) { object_free((object *) new); mcu_mesh_drop_connectivity(mr->m, g); break; }
        for (int e = 0; e <= nel; e++) new->ccs.cptr[e] = e * (g + 1);
        mesh_setconnectivityelement(mesh, 0, (unsigned) g, new);               /* mesh.c:483-485 */
        mesh_link(mesh, (object *) new);
        mesh_freezeconnectivity(mesh);
        mr->conn[g].s = new; mr->conn[g].rix = new->ccs.rix; mr->conn[g].nel = nel; mr->conn[g].ok = true;   /* already on the device */
        mc_stat_devops++;
        return MORPHO_NIL;
    }
    return orig_Mesh_addgrade(v, nargs, args);
}

static builtinfunction orig_Selection_constructor = NULL;
static value mc_Selection_constructor(vm *v, int nargs, value *args) {
    value boundary = MORPHO_FALSE;
    int nfixed = nargs;
    while (mc_ready) {
        builtin_options(v, nargs, args, &nfixed, 1, mc_boundaryoption, &boundary);
        if (nfixed != 1 || !MORPHO_ISTRUE(boundary) || !MORPHO_ISMESH(MORPHO_GETARG(args, 0))) break;       /* not a boundary selection */
        objectmesh *mesh = MORPHO_GETMESH(MORPHO_GETARG(args, 0));
        int maxg = (int) mesh_maxgrade(mesh);
        if (maxg < 2 || maxg > 3) break;                       /* line meshes stay on the reference */
        objectsparse *top = mesh_getconnectivityelement(mesh, 0, (unsigned) maxg), *fac = mesh_getconnectivityelement(mesh, 0, (unsigned) (maxg - 1));
        if (!top || !fac || !sparse_checkformat(top, SPARSE_CCS, true, false) || top->ccs.ncols < mc_connect_min) { mc_log("boundary on the reference:", "facet grade absent or mesh below MORPHO_CUDA_CONNECT_MIN"); break; }
        if (mesh_has_symmetries(mesh, (grade) maxg)) break;
        mc_mirror *mr = mirror_get(mesh, (grade) maxg);
        if (!mr || !mr->conn[maxg].ok) { mc_log("boundary on the reference:", "no mirror of the top grade"); break; }
        mr = mirror_get(mesh, (grade) (maxg - 1));
        if (!mr || !mr->conn[maxg - 1].ok) { mc_log("boundary on the reference:", "no mirror of the facet grade"); break; }
        int bg = 0, n = 0;
        if (mcu_mesh_boundary(mr->m, &bg, &n, NULL, 0) != MCU_OK || bg != maxg - 1) { mc_log("boundary not accelerated:", mcu_last_error()); break; }
        int *ids = malloc(sizeof(int) * (size_t) (n > 0 ? n : 1));
        if (!ids) break;
        if (mcu_mesh_boundary(mr->m, &bg, &n, ids, n) != MCU_OK) { free(ids); break; }
        objectselection *new = object_newselection(mesh);
        if (!new) { free(ids); break; }
        for (int i = 0; i < n; i++) selection_selectelement(new, (grade) bg, ids[i]);     /* ascending ids: the reference's insertion order */
        free(ids);
        selection_addgradelower(new, 0);
        mc_stat_devops++;
        value out = MORPHO_OBJECT(new);
        morpho_bindobjects(v, 1, &out);
        return out;
    }
    return orig_Selection_constructor(v, nargs, args);
}

/* conn(row, col), 1 <= row < col: the grade-lowering matrix (mesh_addlowermatrix, mesh.c:621-660) built on the device and
 * installed where mesh_addconnectivityelement (mesh.c:663-685) would put it; everything that asks for it afterwards —
 * Mesh.connectivitymatrix, selection_addgradelower (selection.c:129-155), the raise matrices (sparse_transpose of this
 * one) — finds it there.  Returns false when nothing was built (already present, mesh too small, grades missing, ...). */
static bool mc_ensure_lower(objectmesh *mesh, int row, int col) {
    if (!mc_ready || row < 1 || col <= row || col > 3 || mesh_getconnectivityelement(mesh, (unsigned) row, (unsigned) col)) return false;
    objectsparse *er = mesh_getconnectivityelement(mesh, 0, (unsigned) row), *ec = mesh_getconnectivityelement(mesh, 0, (unsigned) col);
    if (!er || !ec || !sparse_checkformat(ec, SPARSE_CCS, true, false) || ec->ccs.ncols < mc_connect_min) return false;
    mc_mirror *mr = mirror_get(mesh, (grade) row);
    if (!mr || !mr->conn[row].ok) return false;
    mr = mirror_get(mesh, (grade) col);
    if (!mr || !mr->conn[col].ok) return false;
    const int ncol = ec->ccs.ncols;
    int ncomb = (col == 3 && row == 1) ? 6 : (col == 3 ? 4 : 3);
    int *cptr = malloc(sizeof(int) * ((size_t) ncol + 1)), *rix = malloc(sizeof(int) * ((size_t) ncol * ncomb + 1));
    int nnz = 0, nrows = 0, ncols = 0;
    bool ok = cptr && rix && mcu_mesh_lower(mr->m, row, col, cptr, rix, &nnz, &nrows, &ncols) == MCU_OK;
    objectsparse *new = ok ? object_newsparse(NULL, NULL) : NULL;
    if (new && sparseccs_resize(&new->ccs, nrows, ncols, (unsigned) nnz, false)) {
        memcpy(new->ccs.cptr, cptr, sizeof(int) * ((size_t) ncols + 1));
        memcpy(new->ccs.rix, rix, sizeof(int) * (size_t) nnz);
        mesh_setconnectivityelement(mesh, (unsigned) row, (unsigned) col, new);      /* mesh.c:655-656 */
        mesh_link(mesh, (object *) new);
        mesh_freezeconnectivity(mesh);
        mc_stat_devops++;
    } else {
        if (!ok) mc_log("lowering matrix on the reference:", mcu_last_error());
        if (new) object_free((object *) new);
        ok = false;
    }
    free(cptr); free(rix);
    return ok;
}

static builtinfunction orig_Mesh_connectivitymatrix = NULL;
static value mc_Mesh_connectivitymatrix(vm *v, int nargs, value *args) {
    if (mc_ready && nargs == 2 && MORPHO_ISINTEGER(MORPHO_GETARG(args, 0)) && MORPHO_ISINTEGER(MORPHO_GETARG(args, 1))) {
        objectmesh *mesh = MORPHO_GETMESH(MORPHO_SELF(args));
        int row = MORPHO_GETINTEGERVALUE(MORPHO_GETARG(args, 0)), col = MORPHO_GETINTEGERVALUE(MORPHO_GETARG(args, 1));
        if (row >= 1 && col >= 1 && row <= 3 && col <= 3 && row != col)
            mc_ensure_lower(mesh, row < col ? row : col, row < col ? col : row);     /* a raise matrix is the transpose of the lower one */
    }
    return orig_Mesh_connectivitymatrix(v, nargs, args);
}

/* Selection.addgrade(g) below the selection's highest grade walks conn(g, i) for every higher grade i
 * (selection_addgradelower, selection.c:129-155) */
static builtinfunction orig_Selection_addgrade = NULL;
static value mc_Selection_addgrade(vm *v, int nargs, value *args) {
    if (mc_ready && nargs > 0 && MORPHO_ISINTEGER(MORPHO_GETARG(args, 0))) {
        objectselection *sel = MORPHO_GETSELECTION(MORPHO_SELF(args));
        int g = MORPHO_GETINTEGERVALUE(MORPHO_GETARG(args, 0));
        if (g >= 1 && sel->mesh) for (int i = g + 1; i <= 3 && i < (int) sel->ngrades + 1; i++) mc_ensure_lower(sel->mesh, g, i);
    }
    return orig_Selection_addgrade(v, nargs, args);
}

static int patch_function(const char *name, builtinfunction ours, builtinfunction *orig) {
    objectstring key = MORPHO_STATICSTRING((char *) name);
    value fv;
    if (!dictionary_get(&builtin_functiontable, MORPHO_OBJECT(&key), &fv) || !MORPHO_ISBUILTINFUNCTION(fv)) {
        fprintf(stderr, "[morphocuda] function %s not found as a single builtin; left alone\n", name);
        return 0;
    }
    objectbuiltinfunction *f = MORPHO_GETBUILTINFUNCTION(fv);
    if (f->function == ours) return 1;
    *orig = f->function;
    f->function = ours;
    return 1;
}

#define PATCH(cls, meth) npatched += patch_method(#cls, #meth, mc_##cls##_##meth, &orig_##cls##_##meth)

/* Everything that is done once per process, whichever of `import morphocuda` / `import functionals` comes first. */
static bool mc_core_done = false;
static void mc_core_initialize(void) {
    if (mc_core_done) return;
    mc_core_done = true;
    const char *dev = getenv("MORPHO_CUDA_DEVICE");
    mc_paranoid = getenv("MORPHO_CUDA_PARANOID") && atoi(getenv("MORPHO_CUDA_PARANOID")) != 0;
    mc_verbose = getenv("MORPHO_CUDA_VERBOSE") && atoi(getenv("MORPHO_CUDA_VERBOSE")) != 0;
    mc_verify = getenv("MORPHO_CUDA_VERIFY") && atoi(getenv("MORPHO_CUDA_VERIFY")) != 0;
    if (mcu_init(dev ? atoi(dev) : 0) != MCU_OK) {
        /* no CPU path of ours exists: without a usable GPU nothing is patched and the reference runs as shipped */
        fprintf(stderr, "[morphocuda] CUDA unavailable (%s): functionals stay on the reference builtins\n", mcu_last_error());
        /* MORPHO_CUDA_DRYHOOKS=1 (tests of the host logic on a machine without a GPU): install the trampolines and the
         * class patch anyway; every veneer sees mc_ready == false and hands over to the reference's code */
        if (!(getenv("MORPHO_CUDA_DRYHOOKS") && atoi(getenv("MORPHO_CUDA_DRYHOOKS")) != 0)) return;
        mc_dry = true;
    }
    int nhooks = 0;
    /* (1) generic pass: flush device-resident arguments before any builtin that is not accelerated */
    const char *rmin = getenv("MORPHO_CUDA_RESIDENT_MIN");
    if (rmin) mc_resident_min = atol(rmin);
    bool resident = mc_resident_min >= 0 && !mc_paranoid && !mc_dry;
    if (!resident) mc_resident_min = 0x7fffffffffffL;
    if (resident) nhooks += hook_everything();
    /* (2) writers: what they invalidate */
    nhooks += hook_method("Matrix", MORPHO_ASSIGN_METHOD, WRAP_MATRIX_SELF, resident ? ACC_ASSIGN : ACC_NONE, true);
    nhooks += hook_writers("Matrix", MORPHO_SETINDEX_METHOD, WRAP_MATRIX_SELF);
    nhooks += hook_method("Matrix", MATRIX_SETCOLUMN_METHOD, WRAP_MATRIX_SELF, resident ? ACC_SETCOLUMN : ACC_NONE, true);
    nhooks += hook_method("Matrix", MATRIX_SETCOLUMN_METHOD_DEPRECATED, WRAP_MATRIX_SELF, resident ? ACC_SETCOLUMN : ACC_NONE, true);
    nhooks += hook_method("Matrix", MORPHO_ACC_METHOD, WRAP_MATRIX_SELF, resident ? ACC_ACC : ACC_NONE, true);
    nhooks += hook_writers("Matrix", MATRIX_RESHAPE_METHOD, WRAP_MATRIX_RESHAPE);
    nhooks += hook_writers("Mesh", MESH_SETVERTEXPOSITION_METHOD, WRAP_MESH_VERTS);
    nhooks += hook_method("Mesh", MESH_SETVERTEXMATRIX_METHOD, WRAP_MESH_VERTS, ACC_NONE, false);      /* swaps a pointer, reads nothing */
    nhooks += hook_writers("Mesh", MESH_ADDGRADE_METHOD, WRAP_MESH_CONN);
    nhooks += hook_writers("Mesh", MESH_REMOVEGRADE_METHOD, WRAP_MESH_CONN);
    nhooks += hook_writers("Mesh", MESH_RESETCONNECTIVITY_METHOD, WRAP_MESH_CONN);
    nhooks += hook_writers("Mesh", MESH_ADDSYMMETRY_METHOD, WRAP_MESH_CONN);
    nhooks += hook_writers("Field", MORPHO_SETINDEX_METHOD, WRAP_FIELD_SELF);
    nhooks += hook_writers("Field", MORPHO_ASSIGN_METHOD, WRAP_FIELD_SELF);
    nhooks += hook_writers("Field", MORPHO_ACC_METHOD, WRAP_FIELD_SELF);
    nhooks += hook_writers("Sparse", SPARSE_SETROWINDICES_METHOD, WRAP_CONN_ALL);
    nhooks += hook_writers("Sparse", MORPHO_SETINDEX_METHOD, WRAP_CONN_ALL);
    /* (3) the Matrix methods of optimize.morpho's line searches, on device-resident matrices (mc_resident.h) */
    if (resident) {
        hook_method("Matrix", MORPHO_ADD_METHOD, WRAP_NONE, ACC_ADD, true);
        hook_method("Matrix", MORPHO_ADDR_METHOD, WRAP_NONE, ACC_ADDR, true);
        hook_method("Matrix", MORPHO_SUB_METHOD, WRAP_NONE, ACC_SUB, true);
        hook_method("Matrix", MORPHO_SUBR_METHOD, WRAP_NONE, ACC_SUBR, true);
        hook_method("Matrix", MORPHO_MUL_METHOD, WRAP_NONE, ACC_MUL, true);
        hook_method("Matrix", MORPHO_MULR_METHOD, WRAP_NONE, ACC_MULR, true);
        hook_method("Matrix", MORPHO_DIV_METHOD, WRAP_NONE, ACC_DIV, true);
        hook_method("Matrix", MORPHO_CLONE_METHOD, WRAP_NONE, ACC_CLONE, true);
        hook_method("Matrix", MATRIX_INNER_METHOD, WRAP_NONE, ACC_INNER, true);
        hook_method("Matrix", MATRIX_NORM_METHOD, WRAP_NONE, ACC_NORM, true);
        hook_method("Matrix", MORPHO_SUM_METHOD, WRAP_NONE, ACC_SUM, true);
        hook_method("Matrix", MATRIX_GETCOLUMN_METHOD, WRAP_NONE, ACC_GETCOLUMN, true);
        hook_method("Matrix", MORPHO_COUNT_METHOD, WRAP_NONE, ACC_NONE, false);            /* metadata only */
        hook_method("Matrix", MATRIX_DIMENSIONS_METHOD, WRAP_NONE, ACC_NONE, false);
        hook_method("Mesh", MESH_VERTEXMATRIX_METHOD, WRAP_NONE, ACC_NONE, false);         /* returns the object */
        hook_method("Mesh", MORPHO_COUNT_METHOD, WRAP_NONE, ACC_NONE, false);
        hook_method("Mesh", MESH_MAXGRADE_METHOD, WRAP_NONE, ACC_NONE, false);
    }

    /* (4) the functionals themselves; the saved originals are the trampolines of (1): flush, then the reference builtin */
    int npatched = 0;
    PATCH(Length, total); PATCH(Length, integrand); PATCH(Length, gradient);
    PATCH(Area, total); PATCH(Area, integrand); PATCH(Area, gradient);
    PATCH(VolumeEnclosed, total); PATCH(VolumeEnclosed, integrand); PATCH(VolumeEnclosed, gradient);
    PATCH(Volume, total); PATCH(Volume, integrand); PATCH(Volume, gradient);
    PATCH(AreaEnclosed, total); PATCH(AreaEnclosed, integrand);
    PATCH(GradSq, total); PATCH(GradSq, integrand); PATCH(GradSq, gradient); PATCH(GradSq, fieldgradient);
    PATCH(NormSq, total); PATCH(NormSq, integrand); PATCH(NormSq, gradient); PATCH(NormSq, fieldgradient);
    PATCH(Nematic, total); PATCH(Nematic, integrand); PATCH(Nematic, gradient); PATCH(Nematic, fieldgradient);
    PATCH(NematicElectric, total); PATCH(NematicElectric, integrand); PATCH(NematicElectric, gradient); PATCH(NematicElectric, fieldgradient);
    PATCH(LineCurvatureSq, total); PATCH(LineCurvatureSq, integrand); PATCH(LineCurvatureSq, gradient);
    PATCH(EquiElement, total); PATCH(EquiElement, integrand); PATCH(EquiElement, gradient);
    PATCH(ScalarPotential, total); PATCH(ScalarPotential, integrand); PATCH(ScalarPotential, gradient);
    PATCH(LineIntegral, total); PATCH(LineIntegral, integrand); PATCH(LineIntegral, gradient); PATCH(LineIntegral, fieldgradient);
    PATCH(AreaIntegral, total); PATCH(AreaIntegral, integrand); PATCH(AreaIntegral, gradient); PATCH(AreaIntegral, fieldgradient);
    PATCH(VolumeIntegral, total); PATCH(VolumeIntegral, integrand); PATCH(VolumeIntegral, gradient); PATCH(VolumeIntegral, fieldgradient);
    /* (5) connectivity construction on the device */
    const char *cmin = getenv("MORPHO_CUDA_CONNECT_MIN");
    if (cmin) mc_connect_min = atol(cmin);
    if (mc_connect_min >= 0) {
        mc_boundaryoption = builtin_internsymbolascstring(SELECTION_BOUNDARYOPTION);
        npatched += patch_method("Mesh", MESH_ADDGRADE_METHOD, mc_Mesh_addgrade, &orig_Mesh_addgrade);
        npatched += patch_function(SELECTION_CLASSNAME, mc_Selection_constructor, &orig_Selection_constructor);
        npatched += patch_method("Mesh", MESH_CONNECTIVITYMATRIX_METHOD, mc_Mesh_connectivitymatrix, &orig_Mesh_connectivitymatrix);
        npatched += patch_method("Selection", SELECTION_ADDGRADEMETHOD, mc_Selection_addgrade, &orig_Selection_addgrade);
    }

    object probe;
    memset(&probe, 0, sizeof(probe));
    probe.type = OBJECT_MESH;
    objecttypedefn *defn = object_getdefn(&probe);
    if (defn && defn->freefn != mc_meshfree) { orig_meshfree = defn->freefn; defn->freefn = mc_meshfree; }
    probe.type = OBJECT_FIELD;
    defn = object_getdefn(&probe);
    if (defn && defn->freefn != mc_fieldfree) { orig_fieldfree = defn->freefn; defn->freefn = mc_fieldfree; }
    probe.type = OBJECT_MATRIX;
    defn = object_getdefn(&probe);
    if (resident && defn && defn->freefn != mc_matrixfree) { orig_matrixfree = defn->freefn; defn->freefn = mc_matrixfree; }

    pairwise_initialize();
    mc_ready = !mc_dry;
    if (mc_verbose) fprintf(stderr, "[morphocuda] ready: %d builtins re-pointed, %d writers hooked\n", npatched, nhooks);
}

void morphocuda_initialize(void) {
    /* the script-visible helpers go into the table of the extension being imported (extensions.c:140-156) */
    builtin_addfunction("cudastats", mc_stats, MORPHO_FN_ALLOCATES);
    builtin_addfunction("cudatranslate", mc_translate_fn, MORPHO_FN_ALLOCATES);
    mc_core_initialize();
}

void morphocuda_finalize(void) {
    if (!mc_core_done) return;
    mc_core_done = false;
    mc_ready = false;
    if (mc_verify) mc_verify_report();
    while (nshadows > 0) shadow_drop(&shadows[nshadows - 1]);
    while (nmirrors > 0) mirror_drop(nmirrors - 1);
    fields_drop_for(NULL);
    mcu_finalize();
}
