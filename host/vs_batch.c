/* vs_batch -- batch driver: many voices in one vs_synth_batch() call over all visible GPUs.
 *
 *   vs_batch -m manifest.txt [-G ngpus]
 *
 * Each manifest line describes one voice with the reference tools' own flags:
 *   <out.wav> <vowel> <seed> [flowgen_shimmer flags, e.g. -d 1 -f 120 -j 1 -s 3] [-V gain] [-P pre]
 * which is what   VS_SEED=<seed> flowgen_shimmer -o tmp.wav <flags>; vowel -i tmp.wav -o <out.wav> -v <vowel>
 * would produce, without the intermediate file. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "voicesynth.h"
#include "vs_cli.h"
#include "vs_wav.h"

#define MAXTOK 64

int main(int argc, char **argv)
{
    const char *manifest = NULL;
    int ngpu = 0;
    for (int i = 1; i + 1 < argc; i += 2) {
        if (!strcmp(argv[i], "-m")) manifest = argv[i + 1];
        else if (!strcmp(argv[i], "-G")) ngpu = atoi(argv[i + 1]);
    }
    if (!manifest) { puts("usage: vs_batch -m manifest.txt [-G ngpus]\nline: <out.wav> <vowel> <seed> [flowgen_shimmer flags] [-V gain] [-P pre]"); return 0; }
    FILE *mf = fopen(manifest, "r");
    if (!mf) { perror(manifest); return 1; }

    size_t n = 0, cap = 0;
    vs_cli_flow *rows = NULL;
    char **outs = NULL;
    uint8_t *preset = NULL;
    uint32_t *seed = NULL;
    float *gain = NULL, *pre = NULL;
    char line[4096];
    while (fgets(line, sizeof line, mf)) {
        char *tok[MAXTOK];
        int nt = 0;
        for (char *t = strtok(line, " \t\r\n"); t && nt < MAXTOK; t = strtok(NULL, " \t\r\n")) tok[nt++] = t;
        if (nt == 0 || tok[0][0] == '#') continue;
        if (nt < 3) { fprintf(stderr, "vs_batch: line %zu: need <out> <vowel> <seed>\n", n + 1); return 1; }
        if (n == cap) {
            cap = cap ? cap * 2 : 1024;
            rows = realloc(rows, cap * sizeof *rows); outs = realloc(outs, cap * sizeof *outs);
            preset = realloc(preset, cap); seed = realloc(seed, cap * sizeof *seed);
            gain = realloc(gain, cap * sizeof *gain); pre = realloc(pre, cap * sizeof *pre);
        }
        outs[n] = strdup(tok[0]);
        preset[n] = (uint8_t)tok[1][0];
        seed[n] = (uint32_t)strtoul(tok[2], NULL, 10);
        gain[n] = 10.0f; pre[n] = 1.0f;
        /* split our two extra flags from the flowgen flags */
        char *fargv[MAXTOK + 1];
        int fargc = 1;
        fargv[0] = "vs_batch";
        for (int i = 3; i < nt; i++) {
            if (!strcmp(tok[i], "-V") && i + 1 < nt) gain[n] = (float)atof(tok[++i]);
            else if (!strcmp(tok[i], "-P") && i + 1 < nt) pre[n] = (float)atof(tok[++i]);
            else fargv[fargc++] = tok[i];
        }
        if (fargc == 1) { fargv[fargc++] = "-d"; fargv[fargc++] = "1"; }
        if (vs_cli_parse_flow(fargc, fargv, 0, &rows[n])) { fprintf(stderr, "vs_batch: line %zu: bad flowgen flags\n", n + 1); return 1; }
        n++;
    }
    fclose(mf);
    if (!n) return 0;

    /* AoS rows -> the SoA the ABI takes */
    float *dur = malloc(n * 4), *jit = malloc(n * 4), *shm = malloc(n * 4), *cq = malloc(n * 4), *K = malloc(n * 4),
          *Kv = malloc(n * 4), *F0 = malloc(n * 4), *DC = malloc(n * 4), *noise = malloc(n * 4);
    int32_t *amp = malloc(n * 4), *fs = malloc(n * 4);
    uint8_t *flags = malloc(n);
    for (size_t i = 0; i < n; i++) {
        dur[i] = rows[i].dur; jit[i] = rows[i].jitter; shm[i] = rows[i].shimmer; cq[i] = rows[i].cq; K[i] = rows[i].K;
        Kv[i] = rows[i].Kvar; F0[i] = rows[i].F0; DC[i] = rows[i].DC; noise[i] = rows[i].noise; amp[i] = rows[i].amp;
        fs[i] = rows[i].fs; flags[i] = rows[i].flags;
    }
    vs_flow_params p = {dur, jit, shm, cq, K, Kv, F0, DC, noise, amp, fs, flags, seed};
    vs_filter_params f = {preset, gain, pre};

    int count = vs_device_count();
    if (ngpu <= 0 || ngpu > count) ngpu = count;
    int devs[16];
    for (int d = 0; d < ngpu && d < 16; d++) devs[d] = d;
    vs_ctx *ctx = NULL;
    int rc = vs_ctx_create(&ctx, devs, ngpu, 0);
    if (rc) { fprintf(stderr, "vs_batch: %s\n", vs_strerror(rc)); return 1; }

    uint64_t *ns = malloc(n * sizeof *ns), *offs = malloc(n * sizeof *offs), total = 0;
    vs_flow_nsamples(&p, n, ns);
    for (size_t i = 0; i < n; i++) { offs[i] = total; total += (ns[i] + 7) & ~7ull; }      /* 16-byte aligned rows */
    int16_t *pcm = vs_host_alloc(total * sizeof *pcm);
    if (!pcm) { fprintf(stderr, "vs_batch: out of pinned memory\n"); return 1; }
    rc = vs_synth_batch(ctx, &p, &f, n, pcm, offs, NULL);
    if (rc) { fprintf(stderr, "vs_batch: %s (%s)\n", vs_strerror(rc), vs_last_error(ctx)); return 1; }
    vs_timing t;
    vs_get_timing(ctx, &t);
    for (size_t i = 0; i < n; i++) {
        FILE *o = fopen(outs[i], "wb");
        if (!o) { perror(outs[i]); return 1; }
        vs_wav_write_header(o, (uint32_t)fs[i], (uint32_t)(ns[i] * 2));
        fwrite(pcm + offs[i], 2, ns[i], o);
        fclose(o);
    }
    printf("vs_batch: %zu voices, %llu samples on %d GPU(s): plan %.3f ms, render %.3f ms, %u launches\n", n,
           (unsigned long long)t.samples, ngpu, t.plan_ms, t.render_ms, t.launches);
    vs_host_free(pcm);
    vs_ctx_destroy(ctx);
    return 0;
}
