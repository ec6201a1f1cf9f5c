/* morpho_oracle.h — CPU ORACLE for the functional hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's serial algorithms (Morpho-lang/morpho
 * 0.6.4, src/geometry/functional.c and friends); every function cites the
 * reference file:line it follows.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load this; the product (libmorpho_cuda.so and
 * morpho_b200/) never does.
 *
 * PARITY PINNING: this restatement is checked (tests/test_oracle_*.py) against
 *   (1) the reference's own golden values in test/functionals/ (6 s.f.), and
 *   (2) raw FP64 outputs of the UNMODIFIED reference compiled from
 *       /root/reference/src (oracle/_ref/, built by oracle/refbuild/build_ref.sh),
 *       committed as fixtures under tests/golden/ together with the script that
 *       generated them (tests/golden/make_golden.py).
 *
 * Third-party arithmetic the reference reaches through un-vendored libraries
 * (cblas_ddot/dnrm2/daxpy, LAPACKE_dgesv; no version pinned by the
 * reference) is restated here as the textbook (netlib reference BLAS/LAPACK)
 * algorithms: differences from any particular optimised BLAS are <= a few ulp.
 */
#ifndef MORPHO_ORACLE_H
#define MORPHO_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* Functional ids — shared numbering with include/morpho_cuda.h (MCU_F_*) */
enum {
    MO_LENGTH = 1,
    MO_AREAENCLOSED = 2,
    MO_AREA = 3,
    MO_VOLUMEENCLOSED = 4,
    MO_VOLUME = 5,
    MO_LINECURVATURESQ = 6,
    MO_GRADSQ = 7,
    MO_NORMSQ = 8,
    MO_NEMATIC = 9,
    MO_NEMATICELECTRIC = 10,
    MO_EQUIELEMENT = 11
};

/* Error codes */
enum {
    MO_OK = 0,
    MO_ERR_ARGS = 1,         /* bad arguments */
    MO_ERR_ELNOTFOUND = 2,   /* FnctlELNtFnd: mesh has no elements of the requested grade */
    MO_ERR_DEGENERATE = 3,   /* an element gradient/integrand returned false (nil result) */
    MO_ERR_VOLENCLZERO = 4,  /* VolEnclZero */
    MO_ERR_UNSUPPORTED = 5
};

typedef struct {
    int dim;              /* 2 or 3 */
    int nv;               /* vertices */
    const double *vert;   /* nv*dim, vertex i at vert[i*dim..] (matrix.c:442-444) */
    int nel[4];           /* number of elements per grade (nel[0] ignored) */
    const int *conn[4];   /* conn[g]: nel[g]*(g+1) vertex ids, ascending inside each element */
} mo_mesh;

typedef struct {
    int functional;       /* MO_* */
    int grade;            /* element grade the functional loops over; <=0 → class default */
    /* Nematic */
    double ksplay, ktwist, kbend, pitch;
    int haspitch;
    /* LineCurvatureSq */
    int integrandonly;
    /* fields: field[0] = primary (director / q), field[1] = secondary (electric potential) */
    int psize[2];
    double *field[2];     /* nv*psize, vertex-based P1 fields; not const: FD perturbs & restores */
    int wrt;              /* fieldgradient: which field (0/1) to differentiate */
    /* EquiElement: optional weights, one per element of the grade compared (a 1 x nel Matrix in the reference) */
    const double *weight;
    int nweight;
} mo_functional;

/* functional_fdstepsize (functional.c:46-58) */
double mo_fdstepsize(double x, int order);

/* default grade looped over by a functional class (functional.c Appendix C of SURVEY.md) */
int mo_default_grade(const mo_mesh *m, const mo_functional *f);

/* number of elements the functional loops over (functional_countelements, functional.c:115-129) */
int mo_count(const mo_mesh *m, const mo_functional *f, int *n);

/* sel == NULL → all elements; otherwise nsel element ids in the order given
 * (the reference iterates a selection in hash-slot order, functional.c:226-229). */
int mo_total(const mo_mesh *m, const mo_functional *f, const int *sel, int nsel, double *out);
int mo_integrand(const mo_mesh *m, const mo_functional *f, const int *sel, int nsel, double *out /* n */);
int mo_gradient(const mo_mesh *m, const mo_functional *f, const int *sel, int nsel, double *out /* nv*dim */);
int mo_fieldgradient(const mo_mesh *m, const mo_functional *f, const int *sel, int nsel, double *out /* nv*psize[wrt] */);

/* Connectivity helpers */
/* cs_transpose-style stable counting-sort transpose (sparse.c:1227-1244; mesh.c:663-685):
 * tptr[nv+1], tidx[nel*(g+1)] = element ids ascending per vertex. */
int mo_transpose(int nv, int nel, int nvper, const int *conn, int *tptr, int *tidx);
/* mesh_addgrade for g=1 (mesh.c:419-489): unique edges in first-seen order.  Returns count. */
int mo_edges_from(int nel, int nvper, const int *conn, int *edges_out /* cap nel*C(nvper,2)*2 */);

/* PairwisePotential with Coulomb closures f(r)=1/r, f'(r)=-1/r^2
 * (modules/functionals.morpho:105-139; examples/thomson/thomson.morpho:26). */
/* kind / params: include/morpho_cuda.h MCU_PW_* (0 Coulomb [a], 1 power law [a,p], 2 Yukawa [a,l], 3 Lennard-Jones [e,s]) */
int mo_pairwise_integrand(int n, int dim, const double *x, int kind, const double *params, double cutoff, double *out /* n */);
int mo_pairwise_gradient(int n, int dim, const double *x, int kind, const double *params, double cutoff, double *out /* n*dim */);
double mo_pairwise_total(int n, int dim, const double *x, int kind, const double *params, double cutoff);
int mo_pairwise_coulomb_integrand(int n, int dim, const double *x, double cutoff, double *out /* n */);
int mo_pairwise_coulomb_gradient(int n, int dim, const double *x, double cutoff, double *out /* n*dim */);

/* contiguous binning used by the reference's threaded map and by our GPU partition
 * (functional_binbounds, functional.c:863-879): bounds[nbins+1] */
void mo_binbounds(int nel, int nbins, int *bounds);

#ifdef __cplusplus
}
#endif
#endif
