/* mc_pairwise.h — PairwisePotential (modules/functionals.morpho:87-141) behind the unchanged script API.
 *
 * PairwisePotential is a SCRIPT class in the reference: an interpreted O(N^2) double loop that allocates Matrix objects
 * per pair.  There is no builtin function pointer to re-point, and shadowing the whole module with an extension named
 * `functionals` (extensions win over modules, compile.c:4761-4768) does not work in this version of Morpho: builtin classes
 * are invisible to `class X is Functional` (compiler_findclass looks only at script classes, compile.c:382-390), and
 * modules/meshgen.morpho:103 subclasses Functional — a shadowing extension breaks `import meshgen`.
 *
 * So the module stays exactly as shipped and its compiled class is patched in place at run time: the first builtin call
 * of a program passes through one of our trampolines (morphocuda.c), which scans the classes the program defines
 * (program.classes, program.h:52) for PairwisePotential — by name, superclass name and its methods — and replaces the VALUES of its `integrand` and `gradient` method-table entries (and of the copies its
 * subclasses inherited, compile.c:4193) by builtin functions; OP_INVOKE dispatches builtin functions found in an
 * instance's class like script ones (vm.c:1296-1310).  `total` stays the module's (`self.integrand(mesh).sum()`,
 * functionals.morpho:51-54: both calls arrive here / at the accelerated Matrix.sum).
 *
 * The veneers translate func(r) and grad(r) from bytecode (mc_translate.h) into postfix programs over r.  Programs that
 * are Laurent polynomials in r with grad = func' (Coulomb 1/r of examples/thomson/thomson.morpho:26, power laws,
 * Lennard-Jones) select the specialised instantiations of the all-pairs kernel (mcu_pairwise.cu); anything else runs as
 * a program.  Closures the translator cannot read (control flow), and everything when no GPU is usable, go to the SAVED
 * script method — the reference's own code.
 */
#ifndef MC_PAIRWISE_H
#define MC_PAIRWISE_H

#define FNL_FUNC_PROPERTY    "func"
#define FNL_GRAD_PROPERTY    "grad"
#define FNL_CUTOFF_PROPERTY  "cutoff"

static long mc_stat_pairwise = 0;

/* A postfix program over r as a Laurent polynomial sum_t a_t r^k_t (exponents -64..64), if it is one. */
#define LP_OFF 64
#define LP_N (2 * LP_OFF + 1)
typedef struct { double c[LP_N]; } fnl_laurent;

static bool lp_isconst(const fnl_laurent *p) { for (int k = 0; k < LP_N; k++) if (k != LP_OFF && p->c[k] != 0.0) return false; return true; }
static int lp_nterms(const fnl_laurent *p, int *only) { int n = 0; for (int k = 0; k < LP_N; k++) if (p->c[k] != 0.0) { n++; if (only) *only = k; } return n; }

static bool lp_mul(const fnl_laurent *a, const fnl_laurent *b, fnl_laurent *out) {
    fnl_laurent r;
    memset(&r, 0, sizeof(r));
    for (int i = 0; i < LP_N; i++) if (a->c[i] != 0.0) for (int j = 0; j < LP_N; j++) if (b->c[j] != 0.0) {
        int k = i + j - LP_OFF;
        if (k < 0 || k >= LP_N) return false;
        r.c[k] += a->c[i] * b->c[j];
    }
    *out = r;
    return true;
}

static bool lp_powi(const fnl_laurent *a, int n, fnl_laurent *out) {
    int only = 0;
    fnl_laurent r;
    memset(&r, 0, sizeof(r));
    if (lp_nterms(a, &only) == 1) {                          /* monomial: any integer power */
        int k = (only - LP_OFF) * n + LP_OFF;
        if (k < 0 || k >= LP_N) return false;
        r.c[k] = pow(a->c[only], (double) n);
        *out = r;
        return true;
    }
    if (n < 0 || n > 8) return false;
    r.c[LP_OFF] = 1.0;
    for (int i = 0; i < n; i++) if (!lp_mul(&r, a, &r)) return false;
    *out = r;
    return true;
}

static bool lp_from_program(const mcu_expr_op *prog, int nops, fnl_laurent *out) {
    fnl_laurent *st = (fnl_laurent *) calloc(16, sizeof(fnl_laurent));
    if (!st) return false;
    int sp = 0;
    bool ok = true;
    for (int i = 0; i < nops && ok; i++) {
        const mcu_expr_op op = prog[i];
        switch (op.op) {
            case MCU_X_CONST: if (sp >= 16) { ok = false; break; } memset(&st[sp], 0, sizeof(fnl_laurent)); st[sp++].c[LP_OFF] = op.c; break;
            case MCU_X_POS: if (sp >= 16 || op.a != 0) { ok = false; break; } memset(&st[sp], 0, sizeof(fnl_laurent)); st[sp++].c[LP_OFF + 1] = 1.0; break;
            case MCU_X_ADD: case MCU_X_SUB:
                if (sp < 2) { ok = false; break; }
                sp--;
                for (int k = 0; k < LP_N; k++) st[sp - 1].c[k] += (op.op == MCU_X_ADD ? st[sp].c[k] : -st[sp].c[k]);
                break;
            case MCU_X_MUL: if (sp < 2) { ok = false; break; } sp--; ok = lp_mul(&st[sp - 1], &st[sp], &st[sp - 1]); break;
            case MCU_X_DIV: {
                if (sp < 2) { ok = false; break; }
                sp--;
                fnl_laurent inv;
                ok = lp_nterms(&st[sp], NULL) == 1 && lp_powi(&st[sp], -1, &inv) && lp_mul(&st[sp - 1], &inv, &st[sp - 1]);
                break;
            }
            case MCU_X_NEG: if (sp < 1) { ok = false; break; } for (int k = 0; k < LP_N; k++) st[sp - 1].c[k] = -st[sp - 1].c[k]; break;
            case MCU_X_POWI: if (sp < 1) { ok = false; break; } ok = lp_powi(&st[sp - 1], op.a, &st[sp - 1]); break;
            case MCU_X_POW: {
                if (sp < 2) { ok = false; break; }
                sp--;
                double e = st[sp].c[LP_OFF];
                ok = lp_isconst(&st[sp]) && e == floor(e) && fabs(e) <= 64 && lp_powi(&st[sp - 1], (int) e, &st[sp - 1]);
                break;
            }
            default: ok = false;
        }
    }
    ok = ok && sp == 1;
    if (ok) *out = st[0];
    free(st);
    return ok;
}

/* The potential as the kernel wants it.  kind < 0: only the VM can evaluate these closures. */
typedef struct { int kind; double params[1 + 2 * 6]; } fnl_potential;

static void pairwise_classify(vm *v, value func, value grad, fnl_potential *p) {
    mcu_expr_op pf[MC_MAXPROG], pg[MC_MAXPROG];
    int nf[3] = {0, 0, 0}, ng[3] = {0, 0, 0}, nprog = 0;
    const char *why = "";
    p->kind = -1;
    if (!mc_translate_pointfn(v, func, 1, pf, MC_MAXPROG, nf, &nprog, &why) || nprog != 1) { mc_log("pairwise potential not translated:", why); return; }
    if (!mc_translate_pointfn(v, grad, 1, pg, MC_MAXPROG, ng, &nprog, &why) || nprog != 1) { mc_log("pairwise derivative not translated:", why); return; }
    mc_stat_translated++;
    fnl_laurent lf, lg;
    if (lp_from_program(pf, nf[0], &lf) && lp_from_program(pg, ng[0], &lg)) {
        /* the family kernels derive f' from f: usable only when the grad closure IS the derivative of func */
        bool same = true;
        double scale = 0.0;
        for (int k = 0; k < LP_N; k++) scale = fmax(scale, fabs(lg.c[k]));
        for (int k = 0; k < LP_N && same; k++) {
            double d = (k + 1 < LP_N ? lf.c[k + 1] : 0.0) * (double) (k + 1 - LP_OFF);
            same = fabs(d - lg.c[k]) <= 1e-14 * scale;
        }
        int nt = lp_nterms(&lf, NULL);
        if (same && nt >= 1 && nt <= 6 && lf.c[LP_OFF] == 0.0) {
            if (nt == 1 && lf.c[LP_OFF - 1] != 0.0) { p->kind = MCU_PW_COULOMB; p->params[0] = lf.c[LP_OFF - 1]; return; }
            p->kind = MCU_PW_LAURENT; p->params[0] = (double) nt;
            for (int k = 0, t = 0; k < LP_N; k++) if (lf.c[k] != 0.0) { p->params[1 + 2 * t] = lf.c[k]; p->params[2 + 2 * t] = (double) (k - LP_OFF); t++; }
            return;
        }
    }
    if (mcu_pairwise_program(nf[0], pf, ng[0], pg) == MCU_OK) { p->kind = MCU_PW_PROGRAM; p->params[0] = 0.0; }
    else mc_log("pairwise program not accepted:", mcu_last_error());
}

/* the patched classes and the script methods they had */
#define PW_MAXCLASSES 16
static struct { objectclass *klass; value integrand, gradient; } pw_classes[PW_MAXCLASSES];
static int pw_nclasses = 0;
static program *pw_scanned = NULL;
static int pw_scanned_count = -1;
static value pw_builtin_integrand = MORPHO_NIL, pw_builtin_gradient = MORPHO_NIL;

/* the script method `self`'s class had before the patch (own entry, or the one of the class it inherits from) */
static value pairwise_original(value self, int mode) {
    if (!MORPHO_ISINSTANCE(self)) return MORPHO_NIL;
    for (objectclass *k = MORPHO_GETINSTANCE(self)->klass; k; k = k->superclass)
        for (int i = 0; i < pw_nclasses; i++) if (pw_classes[i].klass == k) return mode == MCU_GRADIENT ? pw_classes[i].gradient : pw_classes[i].integrand;
    return MORPHO_NIL;
}

static value pairwise_eval(vm *v, int nargs, value *args, int mode) {
    value self = MORPHO_SELF(args), ret = MORPHO_NIL;
    value orig = pairwise_original(self, mode);
    if (MORPHO_ISNIL(orig)) { morpho_runtimeerror(v, VM_INVALIDARGS, 1, nargs); return MORPHO_NIL; }
    if (mc_ready && nargs == 1 && MORPHO_ISMESH(MORPHO_GETARG(args, 0))) {
        objectinstance *inst = MORPHO_GETINSTANCE(self);
        objectmesh *mesh = MORPHO_GETMESH(MORPHO_GETARG(args, 0));
        value func = MORPHO_NIL, grad = MORPHO_NIL, cut = MORPHO_NIL;
        prop(inst, FNL_FUNC_PROPERTY, &func); prop(inst, FNL_GRAD_PROPERTY, &grad); prop(inst, FNL_CUTOFF_PROPERTY, &cut);
        double cutoff = 0.0;
        /* `if (self.cutoff && r>self.cutoff) continue`: nil and false switch the test off; a cutoff <= 0 excludes every pair */
        bool hascut = !MORPHO_ISNIL(cut) && !MORPHO_ISFALSE(cut);
        bool device = mesh->vert && mesh->vert->ncols > 0 && (!hascut || (morpho_valuetofloat(cut, &cutoff) && cutoff > 0.0));
        fnl_potential p;
        p.kind = -1;
        mc_mirror *mr = NULL;
        if (device) { pairwise_classify(v, func, grad, &p); device = p.kind >= 0; }
        if (device) { mr = mirror_get(mesh, 0); device = mr != NULL; }
        if (device) {
            const int n = mr->nv, dim = mr->dim;
            const MatrixIdx_t nr = (mode == MCU_GRADIENT) ? dim : n, nc = (mode == MCU_GRADIENT) ? n : 1;   /* Matrix(dim, nv) / Matrix(nv) */
            double *dx = mcu_mesh_vertices_device(mr->m);
            int rc;
            objectmatrix *new = matrix_new(nr, nc, false);
            if (!new) { morpho_runtimeerror(v, ERROR_ALLOCATIONFAILED); return MORPHO_NIL; }
            mc_shadow *s = (shadow_eligible(new) && !mc_paranoid) ? shadow_create(new, false, true) : NULL;
            if (s) rc = mcu_pairwise_device(n, dim, dx, p.kind, p.params, hascut ? cutoff : 0.0, mode, mcu_matrix_device(s->d));
            else {
                mcu_matrix *tmp = mcu_matrix_create(nr, nc);
                rc = tmp ? mcu_pairwise_device(n, dim, dx, p.kind, p.params, hascut ? cutoff : 0.0, mode, mcu_matrix_device(tmp)) : MCU_ERR_ALLOC;
                if (rc == MCU_OK) rc = mcu_matrix_download(tmp, new->elements);
                if (tmp) mcu_matrix_destroy(tmp);
            }
            if (rc == MCU_OK) {
                mc_stat_evals++; mc_stat_pairwise++;
                if (s) mc_stat_devops++;
                value ours = bind_result(v, new);
                if (!mc_verify) return ours;
                /* self-check mode (morphocuda.c): the module's own loop on the same inputs, its result returned */
                int handle = morpho_retainobjects(v, 1, &ours);
                shadow_flush_args(nargs, args);
                bool ok = morpho_invoke(v, self, orig, nargs, args + 1, &ret);
                if (ok) mc_verify_record("PairwisePotential", mode == MCU_GRADIENT ? "gradient" : "integrand", ours, ret);
                morpho_releaseobjects(v, handle);
                return ok ? ret : MORPHO_NIL;
            }
            if (s) shadow_drop(shadow_find(new));
            object_free((object *) new);
            mc_log("pairwise not accelerated:", mcu_last_error());
        }
    }
    /* the module's own method */
    mc_stat_fallbacks++;
    shadow_flush_args(nargs, args);
    if (!morpho_invoke(v, self, orig, nargs, args + 1, &ret)) return MORPHO_NIL;
    return ret;
}

static value Pairwise_integrand(vm *v, int nargs, value *args) { return pairwise_eval(v, nargs, args, MCU_INTEGRAND); }
static value Pairwise_gradient(vm *v, int nargs, value *args) { return pairwise_eval(v, nargs, args, MCU_GRADIENT); }

static bool pw_namedclass(objectclass *k, const char *name) { return k && MORPHO_ISSTRING(k->name) && !strcmp(MORPHO_GETCSTRING(k->name), name); }

static value *pw_methodslot(objectclass *k, const char *name) {
    for (unsigned int i = 0; i < k->methods.capacity; i++) {
        value key = k->methods.contents[i].key;
        if (MORPHO_ISSTRING(key) && !strcmp(MORPHO_GETCSTRING(key), name)) return &k->methods.contents[i].val;
    }
    return NULL;
}

/* Called from the trampolines: patch the PairwisePotential of the program `v` runs, once per program (and again when
 * the program has grown: the interactive CLI compiles line by line into one program). */
static void pairwise_scan(vm *v) {
    program *p = v->current;
    if (!p || MORPHO_ISNIL(pw_builtin_integrand)) return;
    if (p == pw_scanned && pw_scanned_count == (int) p->classes.count) return;
    pw_scanned = p; pw_scanned_count = (int) p->classes.count;
    for (unsigned int c = 0; c < p->classes.count; c++) {
        if (!MORPHO_ISCLASS(p->classes.data[c])) continue;
        objectclass *k = MORPHO_GETCLASS(p->classes.data[c]);
        if (!pw_namedclass(k, "PairwisePotential") || !pw_namedclass(k->superclass, "Functional")) continue;
        value *init = pw_methodslot(k, MORPHO_INITIALIZER_METHOD), *fi = pw_methodslot(k, FUNCTIONAL_INTEGRAND_METHOD), *fg = pw_methodslot(k, FUNCTIONAL_GRADIENT_METHOD);
        /* the module's class: init (a metafunction because of the optional cutoff), buildtree, script integrand(mesh) and
         * gradient(mesh); builtin values here mean "patched already" */
        if (!init || !fi || !fg || !pw_methodslot(k, "buildtree") || !MORPHO_ISFUNCTION(*fi) || !MORPHO_ISFUNCTION(*fg)) continue;
        if (MORPHO_GETFUNCTION(*fi)->nargs != 1 || MORPHO_GETFUNCTION(*fg)->nargs != 1) continue;
        if (pw_nclasses == PW_MAXCLASSES) return;
        value oi = *fi, og = *fg;
        pw_classes[pw_nclasses].klass = k; pw_classes[pw_nclasses].integrand = oi; pw_classes[pw_nclasses].gradient = og;
        pw_nclasses++;
        /* the class itself and every class that inherited these very functions (dictionary_copy at compile.c:4193) */
        for (unsigned int d = 0; d < p->classes.count; d++) {
            if (!MORPHO_ISCLASS(p->classes.data[d])) continue;
            objectclass *kd = MORPHO_GETCLASS(p->classes.data[d]);
            value *si = pw_methodslot(kd, FUNCTIONAL_INTEGRAND_METHOD), *sg = pw_methodslot(kd, FUNCTIONAL_GRADIENT_METHOD);
            if (si && MORPHO_ISSAME(*si, oi)) *si = pw_builtin_integrand;
            if (sg && MORPHO_ISSAME(*sg, og)) *sg = pw_builtin_gradient;
        }
        mc_log("PairwisePotential patched", "");
    }
}

/* once per process (mc_core_initialize): the two builtin functions that take the script methods' places */
static void pairwise_initialize(void) {
    pw_builtin_integrand = builtin_addfunction("__cuda_pairwise_integrand", Pairwise_integrand, MORPHO_FN_REENTRANT);
    pw_builtin_gradient = builtin_addfunction("__cuda_pairwise_gradient", Pairwise_gradient, MORPHO_FN_REENTRANT);
}

#endif
