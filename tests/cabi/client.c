/* A plain-C client of libtiray.so: proves include/tiray.h is valid C99 and that the boundary needs nothing but pointers and
 * sizes.  Parses an OBJ through tr_obj_*, prints the material summary, then asks for a device context.
 * exit codes: 0 = context created and destroyed, 3 = no CUDA device (TR_ERR_NO_DEVICE: there is no CPU fallback), 1 = error */
#include <stdio.h>
#include <stdlib.h>
#include "tiray.h"

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: client file.obj\n"); return 1; }
    tr_obj* obj = NULL;
    if (tr_obj_open(argv[1], &obj) != TR_OK) { fprintf(stderr, "obj: %s\n", tr_obj_last_error()); return 1; }
    int nm = tr_obj_material_count(obj);
    long long total = 0;
    for (int k = 0; k < nm; ++k) {
        char name[64]; double props[9]; int64_t nv = 0; int has_vt = 0, has_vn = 0;
        if (tr_obj_material(obj, k, name, (int)sizeof(name), props, &nv, &has_vt, &has_vn) != TR_OK) return 1;
        double* rows = (double*)malloc((size_t)(nv > 0 ? nv : 1) * 9 * sizeof(double));
        if (nv > 0 && tr_obj_material_vertices(obj, k, rows) != TR_OK) return 1;
        double sx = 0.0; for (int64_t i = 0; i < nv; ++i) sx += rows[i * 9];
        printf("material %d %s tris %lld Kd %.3f %.3f %.3f Ke %.3f d %.3f Ns %.3f Ni %.3f vt %d vn %d sumx %.6f\n", k, name, (long long)(nv / 3),
               props[0], props[1], props[2], props[3], props[6], props[7], props[8], has_vt, has_vn, sx);
        total += nv / 3;
        free(rows);
    }
    tr_obj_close(obj);
    printf("triangles %lld devices %d\n", total, tr_device_count());
    tr_ctx* ctx = NULL;
    int rc = tr_ctx_create(0, &ctx);
    if (rc == TR_ERR_NO_DEVICE) { printf("no device: %s\n", tr_last_error(NULL)); return 3; }
    if (rc != TR_OK) { fprintf(stderr, "ctx: %s\n", tr_last_error(NULL)); return 1; }
    tr_stats st;
    if (tr_film_create(ctx, 64, 64) != TR_OK || tr_stats_get(ctx, &st) != TR_OK) { fprintf(stderr, "%s\n", tr_last_error(ctx)); return 1; }
    printf("context ok\n");
    tr_ctx_destroy(ctx);
    return 0;
}
