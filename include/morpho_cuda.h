/* morpho_cuda.h — C ABI of libmorpho_cuda.so ("libmorpho-cuda").
 *
 * The B200-native implementation of ONE hot path of Morpho (Morpho-lang/morpho
 * 0.6.4): the per-element integrand / total / gradient / fieldgradient maps of
 * the geometric functionals and the scatter of per-element gradients into the
 * vertex-gradient Matrix.  Plain pointers and sizes only; no Morpho, torch or
 * C++ types cross this boundary.  All entry points return an int status
 * (MCU_OK == 0); no exception or longjmp crosses the ABI.  The library is
 * single-caller (Morpho builtins run on the interpreter thread, one call at a
 * time per VM: src/builtin/builtin.h:54, src/core/vm.c:1300-1308) and owns its
 * CUDA stream.  There is NO CPU fallback: every compute entry point fails with
 * MCU_ERR_CUDA when no sm_100 device is usable.
 *
 * Each entry point cites the reference interface it replaces.  The
 * reference-side binding (a Morpho extension that patches the builtin method
 * tables) is morpho_b200/ext/morphocuda.c; see INTEGRATION.md.
 *
 * Data layouts are the reference's own:
 *   vertices : double[nv*dim], vertex i at [i*dim .. i*dim+dim)   (matrix.h:59-73)
 *   elements : int[nel*(g+1)], ids ascending inside an element    (sparse.c:505-510)
 *   gradient : double[nv*dim] same layout as vertices, zero where untouched
 *   integrand: double[nel]    (1 x nel Matrix), zero for unselected elements
 *   field    : double[nv*psize], vertex i at [i*psize ..)         (field.c:419-434)
 */
#ifndef MORPHO_CUDA_H
#define MORPHO_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCU_ABI_VERSION 1

/* ---- status codes; mcu_error_id() maps them to the reference's error ids ---- */
enum {
    MCU_OK = 0,
    MCU_ERR_ARGS = 1,        /* bad argument combination → "FnctlArgs" / "FnctlIntMsh" (functional.h:97-120) */
    MCU_ERR_ELNOTFOUND = 2,  /* "FnctlELNtFnd": mesh has no elements of that grade (functional.c:115-129) */
    MCU_ERR_DEGENERATE = 3,  /* an element gradient returned false → method result is nil (functional.c:392,412-415) */
    MCU_ERR_VOLENCLZERO = 4, /* "VolEnclZero" (functional.c:2148-2151) */
    MCU_ERR_UNSUPPORTED = 5, /* functional/mode not accelerated: caller keeps the reference builtin */
    MCU_ERR_CUDA = 6,        /* CUDA runtime failure or no usable device */
    MCU_ERR_ALLOC = 7        /* "Alloc" */
};

/* ---- functional classes (functional.c:6260-6280); numbering shared with the oracle ---- */
enum {
    MCU_F_LENGTH = 1,          /* functional.c:1948-1993 */
    MCU_F_AREAENCLOSED = 2,    /* functional.c:2000-2056 */
    MCU_F_AREA = 3,            /* functional.c:2063-2124 */
    MCU_F_VOLUMEENCLOSED = 4,  /* functional.c:2131-2178 */
    MCU_F_VOLUME = 5,          /* functional.c:2185-2246 */
    MCU_F_LINECURVATURESQ = 6, /* functional.c:2940-3058 */
    MCU_F_GRADSQ = 7,          /* functional.c:3485-3732 */
    MCU_F_NORMSQ = 8,          /* functional.c:4149-4194 */
    MCU_F_NEMATIC = 9,         /* functional.c:3750-3984 */
    MCU_F_NEMATICELECTRIC = 10,/* functional.c:3996-4142 */
    MCU_F_EQUIELEMENT = 11     /* functional.c:2775-2922 */
};

/* ---- evaluation modes = the builtin method names (functional.h:55-60) ---- */
enum {
    MCU_TOTAL = 0,         /* .total()         → 1 double                         */
    MCU_INTEGRAND = 1,     /* .integrand()     → nel doubles                      */
    MCU_GRADIENT = 2,      /* .gradient()      → nv*dim doubles                   */
    MCU_FIELDGRADIENT = 3  /* .fieldgradient() → nv*psize doubles                 */
};

/* ---- gradient assembly strategies (mcu_set_option("gradient_path", …)) ---- */
enum {
    MCU_PATH_AUTO = 0,
    MCU_PATH_GATHER = 1,   /* CSR-transpose vertex gather, reference summation order */
    MCU_PATH_PATCH = 2,    /* shared-memory patch kernel (triangles; analytic gradients) */
    MCU_PATH_BLOCK = 3     /* shared-memory block kernel (any grade, selections; analytic gradients) */
};

typedef struct mcu_mesh mcu_mesh;           /* device mirror of objectmesh      (mesh.h:25-31)      */
typedef struct mcu_field mcu_field;         /* device mirror of objectfield     (field.h:25-41)     */
typedef struct mcu_selection mcu_selection; /* device mirror of objectselection (selection.h:22-31) */
typedef struct mcu_matrix mcu_matrix;       /* device-resident objectmatrix         (matrix.h:65-73)    */

/* Call descriptor: the part of functional_mapinfo (functional.h:226-241) plus the
 * per-class reference structs (curvatureref functional.c:2935-2939, fieldref 3436-3439,
 * nematicref 3743-3748, nematicelectricref 3996-4000) that the element maps read. */
typedef struct mcu_functional {
    int functional;      /* MCU_F_*                                                    */
    int grade;           /* <=0: class default (GradSq/Nematic*: mesh max grade)       */
    double ksplay, ktwist, kbend, pitch; /* Nematic (functional.c:3757-3791)            */
    int haspitch;
    int integrandonly;   /* LineCurvatureSq (functional.c:2956-2958)                   */
    mcu_field *field[2]; /* [0] director / q ; [1] electric potential (scalar)         */
    int wrt;             /* MCU_FIELDGRADIENT: which of field[] is differentiated      */
    /* EquiElement (equielementref, functional.c:2766-2772): `grade` above is the grade of the elements whose sizes are
     * compared (< 0 or > max grade: the highest grade present); weights are one double per element of that grade in
     * HOST memory (the 1 x nel Matrix of the `weight` option), NULL / 0 for the unweighted functional */
    const double *host_weight;
    int nweight;
} mcu_functional;

/* ---- library lifetime (replaces nothing: extension _initialize/_finalize, extensions.c:134-154) ---- */
int mcu_abi_version(void);
int mcu_init(int device);                   /* select device, create the stream; idempotent */
int mcu_finalize(void);
const char *mcu_last_error(void);           /* human readable text of the last failure */
const char *mcu_error_id(int status);       /* Morpho error id ("VolEnclZero", …) or "" */
int mcu_set_option(const char *key, long value);
long mcu_get_option(const char *key);
int mcu_sync(void);                         /* wait for the library stream (per-mesh status: mcu_mesh_status) */
int mcu_set_stream(void *cuda_stream);      /* adopt a caller-owned cudaStream_t (NULL → own stream) */

/* ---- Mesh mirror: object_newmesh (mesh.c:74), Mesh.setvertexmatrix (mesh.c:1145-1160),
 *      mesh_getconnectivityelement(m,0,g) (mesh.c:296-307) ---- */
mcu_mesh *mcu_mesh_create(int dim, int nv);
int mcu_mesh_destroy(mcu_mesh *m);
int mcu_mesh_set_vertices(mcu_mesh *m, const double *host_vert);      /* H2D of the whole vertex matrix */
int mcu_mesh_set_vertices_device(mcu_mesh *m, const double *dev_vert); /* D2D */
int mcu_mesh_get_vertices(mcu_mesh *m, double *host_vert);            /* D2H */
double *mcu_mesh_vertices_device(mcu_mesh *m);                        /* device pointer of the mirror */
int mcu_mesh_set_connectivity(mcu_mesh *m, int grade, int nel, const int *host_rix);
int mcu_mesh_drop_connectivity(mcu_mesh *m, int grade);                /* Mesh.removegrade (mesh_removegrade, mesh.c:489-498) */
int mcu_mesh_counts(const mcu_mesh *m, int *dim, int *nv, int nel[4]);
/* Mesh.addgrade (mesh_addgrade, mesh.c:419-489): elements of grade g (1 or 2) from the next higher grade present, ids in
 * the reference's order of first discovery; built on the device (sort + scan).  mcu_mesh_get_connectivity copies the
 * incidence (nel * (g+1) ints) of any grade back. */
int mcu_mesh_addgrade(mcu_mesh *m, int grade, int *nel_out);
/* Selection(mesh, boundary=true) (selection_selectboundary, selection.c:211-232): ids, ascending, of the elements of
 * grade maxgrade-1 contained in exactly one element of the top grade (for a line mesh: vertices with one incident line).
 * *grade_out = that grade, *n_out = their number; at most cap ids are copied to host_ids (may be NULL to count). */
int mcu_mesh_boundary(mcu_mesh *m, int *grade_out, int *n_out, int *host_ids, int cap);
int mcu_mesh_get_connectivity(mcu_mesh *m, int grade, int *host_rix);
/* conn(row, col), 1 <= row < col <= 3, the grade-LOWERING matrix (mesh_addlowermatrix, mesh.c:621-660): for every element
 * of grade col the ids, ascending, of the elements of grade row all of whose vertices it contains, in compressed
 * columns.  host_cptr: nel(col) + 1 ints; host_rix: room for nel(col) * C(col+1, row+1) ints; *nnz_out = entries
 * written; *nrows_out / *ncols_out = largest row / column index with an entry, plus one (the size the reference's
 * dictionary-of-keys matrix grows to: trailing elements without entries are not part of it).  The
 * grade-RAISING matrices are transposes (mesh_addconnectivityelement, mesh.c:663-685: sparse_transpose).
 * MCU_ERR_UNSUPPORTED when two elements of grade row have the same vertices. */
int mcu_mesh_lower(mcu_mesh *m, int row, int col, int *host_cptr, int *host_rix, int *nnz_out, int *nrows_out, int *ncols_out);
/* vertex→element incidence as the reference builds it with sparse_transpose → cs_transpose
 * (sparse.c:1227-1244, mesh.c:663-685): tptr[nv+1], tidx[nel*(g+1)] element ids ascending per vertex. */
int mcu_mesh_get_transpose(mcu_mesh *m, int grade, int *host_tptr, int *host_tidx);

/* ---- Field mirror (vertex based, P1): object_newfield (field.c:107-170) ---- */
mcu_field *mcu_field_create(mcu_mesh *m, int psize);
int mcu_field_destroy(mcu_field *f);
int mcu_field_set(mcu_field *f, const double *host_data);
int mcu_field_set_device(mcu_field *f, const double *dev_data);
int mcu_field_get(mcu_field *f, double *host_data);
double *mcu_field_device(mcu_field *f);

/* ---- Selection mirror: element ids of one grade (selection.h:22-31; the integer keys of
 *      sel->selected[g], functional.c:226-229). Order of ids is irrelevant to the result set. ---- */
mcu_selection *mcu_selection_create(mcu_mesh *m, int grade, int n, const int *host_ids);
int mcu_selection_destroy(mcu_selection *s);
int mcu_selection_get_ids(mcu_selection *s, int *host_ids_sorted, int cap); /* returns count */

/* ---- Evaluation = functional_sumintegrand / functional_mapintegrand / functional_mapgradient /
 *      functional_mapnumericalgradient / functional_mapnumericalfieldgradient
 *      (functional.c:962-996, 1049-1085, 1092-1135, 1193-1247, 1428-1503).
 *      mcu_eval copies the result to host memory (the objectmatrix / objectfield the builtin
 *      returns); mcu_eval_device leaves it in device memory and does not synchronise. ---- */
int mcu_eval(mcu_mesh *m, const mcu_functional *f, mcu_selection *sel, int mode, double *host_out);
int mcu_eval_device(mcu_mesh *m, const mcu_functional *f, mcu_selection *sel, int mode, double *dev_out);
/* Status contract.  mcu_eval clears the mesh's status word, evaluates, and returns MCU_ERR_DEGENERATE /
 * MCU_ERR_VOLENCLZERO for THAT evaluation.  mcu_eval_device is asynchronous and cannot: mcu_mesh_status waits for
 * the stream, then reports and clears whatever the device-path evaluations on this mesh raised since the last
 * mcu_mesh_status / mcu_eval (a failed evaluation leaves inf/NaN rows; the reference returns nil, functional.c:392). */
int mcu_mesh_status(mcu_mesh *m);
/* number of doubles `mode` produces for this functional on this mesh */
long mcu_result_size(mcu_mesh *m, const mcu_functional *f, int mode);

/* ---- Device-resident Matrix objects and the vector operations of the optimisers.
 *      optimize.morpho's line search is `vsave = x.clone(); x.acc(-s, force); totals; mesh.setvertexmatrix(vsave)` and
 *      its iterations combine gradients with inner / norm / acc / scalar multiples / column writes
 *      (optimize.morpho:229-303, 461-486, 686-693); the reference runs them as cblas_daxpy / ddot / dnrm2 / dscal on host
 *      arrays.  With these entry points a matrix lives on the device for its whole life: the extension
 *      keeps a lazily synchronised host copy only for readers it does not accelerate.  Layout: column-major
 *      nrows x ncols doubles, as objectmatrix.elements.  Reductions are deterministic (fixed chunks, compensated sums).
 *        mcu_matrix_axpby   out <- sa a + sb b (b NULL: sa a)   matrix_axpy / Matrix +,- / matrix_scale  matrix.c:495-500, 526-529, 1098-1160
 *        mcu_matrix_xpa     out <- a sgn + beta                 matrix_addscalar                          matrix.c:557-565
 *        mcu_matrix_copy    dst <- src                          matrix_copy / matrix_clone                matrix.c:310-320, 503-509
 *        mcu_matrix_inner / norm / sum                          matrix_inner / matrix_norm / matrix_sum   matrix.c:586-604, 621-626
 *        mcu_matrix_set_column / get_column / zero_columns      matrix_setcolumn / getcolumn; Optimizer.fixgrad  matrix.c:447-463
 *      mcu_mesh_bind_vertices makes a mesh read its vertices from such a matrix (Mesh.setvertexmatrix, mesh.c:1145-1160)
 *      without a copy; mcu_eval_matrix evaluates a gradient / integrand straight into one. ---- */
mcu_matrix *mcu_matrix_create(int nrows, int ncols);
int mcu_matrix_destroy(mcu_matrix *m);
int mcu_matrix_dims(const mcu_matrix *m, int *nrows, int *ncols);
double *mcu_matrix_device(mcu_matrix *m);
int mcu_matrix_upload(mcu_matrix *m, const double *host);
int mcu_matrix_download(mcu_matrix *m, double *host);
int mcu_matrix_copy(mcu_matrix *dst, const mcu_matrix *src);
int mcu_matrix_zero(mcu_matrix *m);
int mcu_matrix_axpby(mcu_matrix *out, const mcu_matrix *a, double sa, const mcu_matrix *b, double sb);
int mcu_matrix_xpa(mcu_matrix *out, const mcu_matrix *a, double sgn, double beta);
int mcu_matrix_inner(mcu_matrix *a, const mcu_matrix *b, double *host_out);
int mcu_matrix_norm(mcu_matrix *a, double *host_out);
int mcu_matrix_sum(mcu_matrix *a, double *host_out);
int mcu_matrix_set_column(mcu_matrix *m, int col, const double *host_col);
int mcu_matrix_get_column(mcu_matrix *m, int col, double *host_col);
int mcu_matrix_zero_columns(mcu_matrix *m, int n, const int *host_cols);
int mcu_mesh_bind_vertices(mcu_mesh *m, mcu_matrix *vert);
int mcu_eval_matrix(mcu_mesh *m, const mcu_functional *f, mcu_selection *sel, int mode, mcu_matrix *out);

/* ---- Line / area / volume integrals of an integrand over the elements (grade 1, 2, 3): integrate_integrate
 *      (integrate.c:769-800), the default element integrator behind LineIntegral / AreaIntegral / VolumeIntegral
 *      (functional.c:5176-5245, 5354-5428, 5471-5546): adaptive Gauss-Kronrod 7/15 on lines (integrate_lineint), nested
 *      Grundmann-Moeller rules with quadrisection on triangles (integrate_areaint), the 35/56-point pair with 8-way
 *      subdivision on tetrahedra (integrate_volint, integrate.c:677-752), accuracy goal 1e-6 relative to the
 *      element's estimate, result multiplied by the element size.  The reference calls a Morpho closure at every node;
 *      here the integrand is a postfix program over the interpolated position (MCU_X_POS a = 0,1,2) and the
 *      interpolated components of the given P1 fields (MCU_X_FIELD a = flattened component index, fields in
 *      argument order).  mode MCU_TOTAL (1 double) or MCU_INTEGRAND (nel doubles, zero for unselected elements). ---- */
enum {
    MCU_X_CONST = 0, MCU_X_POS = 1, MCU_X_FIELD = 2,
    MCU_X_ADD = 3, MCU_X_SUB = 4, MCU_X_MUL = 5, MCU_X_DIV = 6,
    MCU_X_NEG = 7, MCU_X_POWI = 8 /* integer power a */, MCU_X_SQRT = 9, MCU_X_SIN = 10, MCU_X_COS = 11, MCU_X_EXP = 12,
    MCU_X_LOG = 13, MCU_X_ABS = 14,
    MCU_X_POW = 15 /* binary: pow(l, r) */, MCU_X_TAN = 16, MCU_X_ASIN = 17, MCU_X_ACOS = 18, MCU_X_ATAN = 19,
    MCU_X_ATAN2 = 20 /* binary: atan2(l, r) */, MCU_X_SINH = 21, MCU_X_COSH = 22, MCU_X_TANH = 23, MCU_X_FLOOR = 24,
    MCU_X_CEIL = 25, MCU_X_LOG10 = 26,
    /* element constants, the helper builtins of functional.c:4396-4745: a = component of the unit tangent of a line
     * element / the unit normal of a triangle in 3D; MCU_X_GRADQ a = 3 * (flattened field component) + k is the
     * derivative along x_k of the P1 interpolant of that component on a triangle (gradsq_evaluategradient) or a
     * tetrahedron in 3D (gradsq_evaluategradient3d) */
    MCU_X_TANGENT = 27, MCU_X_NORMAL = 28, MCU_X_GRADQ = 29,
    /* control flow of a closure, branch free: comparisons give 1.0 / 0.0 with the VM's notion of equal floats
     * (morpho_doubleeqtest, value.c:48-54: equal within DBL_EPSILON relative; vm.c:1095-1135), NOT of such a value, and
     * SELECT pops (condition, then, else) — both sides of an `if` are evaluated, the condition picks one */
    MCU_X_LT = 30, MCU_X_LE = 31, MCU_X_EQ = 32, MCU_X_NE = 33, MCU_X_NOT = 34, MCU_X_SELECT = 35,
    /* common subexpressions: STORE copies the top of the stack into register a (0..31) and leaves it there, LOAD pushes
     * register a; a register is stored before it is loaded */
    MCU_X_STORE = 36, MCU_X_LOAD = 37
};
typedef struct mcu_expr_op { int op; int a; double c; } mcu_expr_op;
int mcu_integral(mcu_mesh *m, int grade, int nops, const mcu_expr_op *prog, int nfields, mcu_field *const *fields,
                 mcu_selection *sel, int mode, double *host_out);
/* LineIntegral/AreaIntegral/VolumeIntegral gradient and fieldgradient: central differences over the quadrature, as the reference
 * (functional_mapnumericalgradient / functional_mapnumericalfieldgradient, functional.c:1142-1247, 1305-1503 with
 * lineintegral_integrand / areaintegral_integrand / volumeintegral_integrand).  wrt < 0: positions (nv * dim doubles); wrt = k: the k-th field
 * (nv * psize_k doubles, the layout of the Field's store). */
int mcu_integral_gradient(mcu_mesh *m, int grade, int nops, const mcu_expr_op *prog, int nfields, mcu_field *const *fields,
                          mcu_selection *sel, int wrt, double *host_out);

/* The same with the reference's RULE-TABLE integrator, which a `method` dictionary selects (integrate(),
 * integrate.c:2849-2866; integrator_configurewithdictionary 2706-2751): `rule` = the "rule" entry (empty: none), `degree` =
 * the "degree" entry (< 0: none), `adapt` = the "adapt" entry (default true).  Rule selection follows
 * integrator_configure (2652-2704): by name, else the lowest rule of at least that degree, else the default of the grade
 * (gauss5, cubtri7, grundmann3d0); the extension chain of the rule (or a second rule of higher degree) gives the error
 * estimate; the worst work item is bisected / quadrisected / cut into eight until the summed error is below 1e-6 of the
 * value (integrator_integrate, 2759-2835).  method == NULL: the default integrator above.  MCU_ERR_UNSUPPORTED when the
 * reference would raise IntgrtrRlNtFnd / IntgrtrRlUnavlb / IntgrtrSbdvsns or an element needs more work items than the
 * device keeps per element (256 / 157 / 102 for lines / triangles / tetrahedra): the caller leaves the call to the reference. */
typedef struct mcu_quad_method { int adapt; int degree; char rule[32]; } mcu_quad_method;
int mcu_integral_method(mcu_mesh *m, int grade, int nops, const mcu_expr_op *prog, int nfields, mcu_field *const *fields,
                        mcu_selection *sel, int mode, const mcu_quad_method *method, double *host_out);
int mcu_integral_gradient_method(mcu_mesh *m, int grade, int nops, const mcu_expr_op *prog, int nfields, mcu_field *const *fields,
                                 mcu_selection *sel, int wrt, const mcu_quad_method *method, double *host_out);

/* ---- ScalarPotential (functional.c:2255-2372): a function of the position evaluated at every (selected) vertex, given
 *      as a postfix program over MCU_X_POS (no field / element ops).  nprog programs are concatenated in progs, program k
 *      has nops[k] ops.  MCU_TOTAL (1 double) and MCU_INTEGRAND (nv doubles): nprog = 1.  MCU_GRADIENT (nv * dim
 *      doubles): nprog = dim gives the components of the gradient function, nprog = 1 takes central differences of the
 *      potential like functional_mapnumericalgradient does when no gradient function was supplied. ---- */
int mcu_pointwise(mcu_mesh *m, int nprog, const int *nops, const mcu_expr_op *progs, mcu_selection *sel, int mode, double *host_out);

/* ---- PairwisePotential (modules/functionals.morpho:87-141), all-pairs FP64.  The reference takes two Morpho closures,
 *      func(r) and grad(r) = func'(r); here the potential is one of these families (the extension recognises them from
 *      the closures' bytecode, anything else arrives as a pair of postfix programs):
 *        MCU_PW_COULOMB  f = a/r                        params = {a}           (NULL: a = 1; examples/thomson/thomson.morpho:26)
 *        MCU_PW_LAURENT  f = sum_t a_t r^k_t, k_t integer  params = {nt, a_0, k_0, a_1, k_1, ...}, nt <= 6
 *                        (power laws, Lennard-Jones 4e((s/r)^12 - (s/r)^6) = {2, 4 e s^12, -12, -4 e s^6, -6}); f' is derived
 *        MCU_PW_YUKAWA   f = a exp(-r/l)/r              params = {a, l}
 *        MCU_PW_PROGRAM  f and f' as postfix programs over r (MCU_X_POS a = 0), set by mcu_pairwise_program
 *      cutoff > 0: pairs with r > cutoff are skipped (functionals.morpho:113,130).
 *      mode MCU_TOTAL / MCU_INTEGRAND (n doubles) / MCU_GRADIENT (n*dim doubles). ---- */
enum { MCU_PW_COULOMB = 0, MCU_PW_LAURENT = 1, MCU_PW_YUKAWA = 2, MCU_PW_PROGRAM = 3 };
int mcu_pairwise_program(int nops_f, const mcu_expr_op *prog_f, int nops_g, const mcu_expr_op *prog_g);
int mcu_pairwise(int n, int dim, const double *host_x, int kind, const double *params, double cutoff,
                 int mode, double *host_out);
int mcu_pairwise_device(int n, int dim, const double *dev_x, int kind, const double *params, double cutoff,
                        int mode, double *dev_out);
/* Row block [row0, row0 + nrows) of the interaction against all n points: the per-GPU share of the row-block partition
 * (SURVEY 8e); outputs are indexed by row - row0; MCU_TOTAL gives the partial energy of the block. */
int mcu_pairwise_rows_device(int n, int dim, const double *dev_x, int kind, const double *params, double cutoff,
                             int mode, int row0, int nrows, double *dev_out);

/* ---- element partition across GPUs: functional_binbounds (functional.c:863-879) ---- */
int mcu_binbounds(int nel, int nbins, int *bounds);

/* ---- multi-GPU (one process per GPU): completion of the gradient rows of vertices shared between ranks.
 *      Elements are partitioned in contiguous id ranges (mcu_binbounds); the reference's analogue is the merge of the
 *      per-thread gradient matrices with matrix_axpy (functional.c:1106-1119).  The exchange runs over NVLink peer
 *      memory (CUDA IPC), not through a collective library: see morpho_b200/csrc/mcu_halo.cu.  At most 16 ranks, 4
 *      gradients (column blocks) per step.
 *        shared_local[n_shared]        local ids of the vertices other ranks also touch, ascending global id
 *        peer_ptr[world+1], peer_pos[] per peer q the positions (in shared_local) of the vertices shared with q,
 *                                      ascending global id — the same order on both sides of a pair
 *        max_m                         longest pair list over all pairs (agreed by all ranks)
 *        contrib_ptr[n_shared+1], contrib_rank[], contrib_k[]
 *                                      per shared vertex the sharing ranks in ascending order (own rank included, k = -1)
 *                                      and the vertex's index k in the (that rank, this rank) pair list
 *      ncomp = doubles exchanged per vertex (number of gradients x dim). ---- */
typedef struct mcu_halo mcu_halo;
mcu_halo *mcu_halo_create(int rank, int world, int ncomp, int n_shared, const int *shared_local, const int *peer_ptr,
                          const int *peer_pos, int max_m, const int *contrib_ptr, const int *contrib_rank,
                          const int *contrib_k);
int mcu_halo_export(mcu_halo *h, unsigned char handle[64]);          /* IPC handle of this rank's receive block */
int mcu_halo_connect(mcu_halo *h, const unsigned char *handles);     /* world x 64 bytes, rank order */
int mcu_halo_exchange(mcu_halo *h, int ngrad, double *const *dev_grads, int dim);   /* asynchronous, library stream */
/* one gradient (column block `slot`) on the library's SECOND stream, underneath the kernels queued on the library stream;
 * mcu_halo_join makes the library stream wait for the exchanges in flight (implied by the next evaluation of the same
 * column block on the attached mesh, by mcu_halo_exchange and by mcu_halo_status) */
int mcu_halo_exchange_async(mcu_halo *h, int slot, double *dev_grad, int dim);
int mcu_halo_join(mcu_halo *h);
/* fused mode: the patch gradient kernel of `m` stores the shared rows into the peers' buffers itself (column block
 * `slot` of the exchanged row, chosen with mcu_halo_bind before each evaluation); mcu_halo_exchange then only raises
 * the flags, waits for the peers and combines */
int mcu_halo_attach(mcu_halo *h, mcu_mesh *m);
int mcu_halo_bind(mcu_halo *h, int slot);
int mcu_halo_status(mcu_halo *h);                                    /* 0 ok, 1 a peer timed out (synchronises) */
int mcu_halo_destroy(mcu_halo *h);

/* ---- measurement helpers (bench.py): CUDA events on the library stream ---- */
int mcu_timer_start(void);
int mcu_timer_stop(float *ms);
long mcu_launch_count(void);                /* kernels launched by this library since mcu_init */
int mcu_device_malloc(void **ptr, size_t bytes);
int mcu_device_free(void *ptr);
int mcu_flush_l2(void);                     /* overwrite a > L2-sized scratch buffer */
/* FP64 peak of the device from a DFMA micro-benchmark (2 flop per DFMA): the roofline denominator of the pairwise kernel */
int mcu_debug_dfma_peak(double *tflops);

/* ---- introspection of the patch planner (pure host code, no CUDA call): used by the CPU tests
 *      of the host logic.  counts = {patches, runs, table words, ints per header, max slots, max table
 *      bytes, halo vertices, alignment slots, modelled shared-load wavefronts, ideal wavefronts}.
 *      Output arrays may be NULL. ---- */
int mcu_debug_patch_plan(int nv, int nel, const int *rix, const double *vert, int max_owned, long counts[10],
                         int *hdr_out, int *runs_out, unsigned *tab_out);

/* the block decomposition of the generic owner-computes gradient kernel (mcu_block.cu), pure host code: counts = {blocks,
 * local vertices, entries, vertices with elements, 1 if every vertex has one}; hdr_out receives 14 ints per block
 * {loc_off, n_local, n_owned, 0, ent_off[8], deg[8] packed in 2 ints}.  eids: selected element ids or NULL. */
int mcu_debug_block_plan(int nv, int dim, int n, int nvper, const int *rix, const int *eids, const double *vert, long counts[5],
                         int *hdr_out, int *loc_out, unsigned *ent_out);

/* timeline diagnostics of the patch kernel (measurement aid): dev_buf receives, per consumer warp of every CTA,
 * {cycles waiting for data, cycles alive}; NULL disarms. */
int mcu_debug_patch_triangles(int nv, int nel, const int *rix, const double *vert, int max_owned,
                              long *n_words_out, int *tri_hdr_out, unsigned *tri_out);
int mcu_debug_patch_profile(long long *dev_buf);
/* the quadrature rules in the order of quadrules[] (integrate.c:1958-1975): info = grade, order, nnodes, index of the
 * extension rule or -1, number of weights; and integrator_configure's choice (rule, error rule or -1) */
int mcu_debug_quadrule(int index, char name[32], int info[5], double *nodes, double *weights, int cap);
int mcu_debug_quadconfigure(int grade, int adapt, int degree, const char *name, int out[2]);

#ifdef __cplusplus
}
#endif
#endif /* MORPHO_CUDA_H */
