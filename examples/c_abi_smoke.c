/* c_abi_smoke.c -- include/rvh.h used from plain C (gcc -std=c11), linked against librvh.so.
 * Without a CUDA device rvh_create must fail with a message (there is no CPU fallback); with one it runs a few steps of
 * the reference-sized scene on synthetic strands and prints the draw-indirect block.
 *   gcc -std=c11 -Wall -Iinclude examples/c_abi_smoke.c -Lrealtime-vulkan-hair_b200 -lrvh -Wl,-rpath,$PWD/realtime-vulkan-hair_b200 -o /tmp/c_abi_smoke */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "rvh.h"

int main(void) {
    rvh_config cfg;
    rvh_default_config(&cfg, 900, 10);                 /* Strand.h:8-9 */
    printf("abi %d, config %zu bytes, rest_length %.6f, grid %d^3, flags %d\n", rvh_abi_version(), sizeof cfg, cfg.rest_length, cfg.grid_dim, cfg.flags);
    float sphere[48], t[3] = { 2.f, 0.f, 1.f }, r[3] = { 0.f, 0.f, 0.f }, s[3] = { 1.f, 1.f, 1.f };
    rvh_collider_build(t, r, s, sphere);               /* Scene.h:28-38 */
    printf("sphere centre %.1f %.1f %.1f, fbm(0.5) %.6f\n", sphere[12], sphere[13], sphere[14], rvh_wind_fbm(0.5f));
    rvh_ctx* ctx = NULL;
    int rc = rvh_create(&ctx, &cfg);
    if (rc != RVH_OK) { printf("rvh_create: %d (%s)\n", rc, rvh_last_error(NULL)); return rc == RVH_ERR_CUDA ? 3 : 1; }
    size_t bytes = (size_t)900 * 48 * 10;
    float* strands = (float*)calloc(bytes / 4, 4);
    for (int k = 0; k < 900; ++k)
        for (int j = 0; j < 10; ++j) {                 /* straight strands above the head, at rest spacing */
            float* p = strands + ((size_t)k * 30 + j) * 4;
            p[0] = -0.5f + 0.001f * k; p[1] = 3.6f + cfg.rest_length * j; p[2] = 0.f; p[3] = 1.f;
        }
    rc = rvh_set_colliders(ctx, sphere, 1);
    if (!rc) rc = rvh_upload_strands_aos(ctx, strands, bytes);
    for (int k = 0; k < 10 && !rc; ++k) rc = rvh_step(ctx, 1.f / 60.f, k / 60.f);
    uint32_t ind[4] = { 0, 0, 0, 0 };
    if (!rc) rc = rvh_download_strands_aos(ctx, strands, bytes);
    if (!rc) rc = rvh_draw_indirect(ctx, ind);
    if (rc) printf("error %d: %s\n", rc, rvh_last_error(ctx));
    else printf("10 steps ok: tip of strand 0 at %.4f %.4f %.4f, draw indirect {%u,%u,%u,%u}, %lld kernel launches\n", strands[36], strands[37], strands[38],
                ind[0], ind[1], ind[2], ind[3], rvh_kernel_launches(ctx));
    rvh_destroy(ctx);
    free(strands);
    return rc ? 1 : 0;
}
