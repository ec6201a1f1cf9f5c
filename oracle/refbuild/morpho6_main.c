/* morpho6_main.c — OUR minimal command-line driver on libmorpho's public C API
 * (src/morpho.h:103-199).  The reference repository builds only the library;
 * its CLI lives in another repository that is not available here.
 * TEST INFRASTRUCTURE ONLY.
 *
 *   morpho6 [-wN] [-D] file.morpho        (-D: print the bytecode listing before running)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "morpho.h"

static char *slurp(const char *path) {
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    char *buf = malloc((size_t) n + 1);
    if (buf && fread(buf, 1, (size_t) n, f) != (size_t) n) { free(buf); buf = NULL; }
    if (buf) buf[n] = '\0';
    fclose(f);
    return buf;
}

int main(int argc, const char *argv[]) {
    int nthreads = 0, disasm = 0; const char *file = NULL;
    for (int i = 1; i < argc; i++) {
        if (strncmp(argv[i], "-w", 2) == 0) nthreads = atoi(argv[i] + 2);
        else if (strcmp(argv[i], "-D") == 0) disasm = 1;
        else file = argv[i];
    }
    if (!file) { fprintf(stderr, "usage: morpho6 [-wN] file.morpho\n"); return 2; }
    char *src = slurp(file);
    if (!src) { fprintf(stderr, "cannot read %s\n", file); return 2; }

    morpho_initialize();
    morpho_setthreadnumber(nthreads);
    morpho_setargs(argc, argv);
    program *p = morpho_newprogram();
    compiler *c = morpho_newcompiler(p);
    vm *v = morpho_newvm();
    error err; error_init(&err);
    int rc = 0;
    if (morpho_compile(src, c, false, &err)) {
        if (disasm) morpho_disassemble(v, p, NULL);
        if (!morpho_run(v, p)) {
            error *e = morpho_geterror(v);
            printf("Error '%s': %s\n", e->id ? e->id : "?", e->msg);
            rc = 1;
        }
    } else {
        printf("Compile error '%s' [line %i]: %s\n", err.id ? err.id : "?", err.line, err.msg);
        rc = 1;
    }
    morpho_freevm(v);
    morpho_freecompiler(c);
    morpho_freeprogram(p);
    morpho_finalize();
    free(src);
    return rc;
}
