/* flowgen_shimmer -- drop-in for the reference tool of the same name: same flags, same 16-bit WAV
 * payload, same per-period stdout lines; the per-period loop (flowgen_shimmer.c:246-423) runs on the
 * GPU through vs_flowgen_batch().  Differences, all deliberate: the header is the canonical 44-byte
 * one on every ABI, -r 22050 is accepted (the reference's range test rejects exactly that value,
 * :537), and failures exit non-zero. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "voicesynth.h"
#include "vs_cli.h"
#include "vs_wav.h"

static int die(vs_ctx *ctx, int rc, const char *what)
{
    fprintf(stderr, "flowgen_shimmer: %s: %s (%s)\n", what, vs_strerror(rc), ctx ? vs_last_error(ctx) : "");
    if (ctx) vs_ctx_destroy(ctx);
    return 1;
}

int main(int argc, char **argv)
{
    vs_cli_flow a;
    if (vs_cli_parse_flow(argc, argv, 1, &a)) { vs_cli_flow_usage(); return 0; }      /* the reference exits 0 here too */

    const uint32_t seed = vs_cli_seed();
    vs_flow_params p;
    memset(&p, 0, sizeof p);
    p.dur = &a.dur; p.jitter = &a.jitter; p.shimmer = &a.shimmer; p.cq = &a.cq; p.K = &a.K; p.Kvar = &a.Kvar;
    p.F0 = &a.F0; p.DC = &a.DC; p.noise = &a.noise; p.amp = &a.amp; p.fs = &a.fs; p.flags = &a.flags; p.seed = &seed;

    uint64_t n = 0, max_periods = 0;
    size_t bad = 0;
    int rc = vs_flow_validate(&p, 1, &bad);
    if (rc) return die(NULL, rc, "parameters");
    vs_flow_nsamples(&p, 1, &n);
    vs_flow_max_periods(&p, 1, &max_periods);

    FILE *out = fopen(a.out_path, "wb");
    if (!out) { printf("Error while creating %s\n", a.out_path); return 1; }
    /* header data size as the reference states it: (long)(dur*fs*2) in float arithmetic (:555) */
    const uint32_t data_bytes = (uint32_t)(long)(a.dur * (float)a.fs * 2);
    if (vs_wav_write_header(out, (uint32_t)a.fs, data_bytes)) { printf("Error while writing header to %s\n", a.out_path); return 1; }

    /* the reference's banner, line for line (flowgen_shimmer.c:576-588; no newline after its second sentence) */
    printf("(c) Maurilio N. Vieira, 1996\nSynthetic vowel generator\n");
    printf("ported to gcc - Joao SANSAO, Feb. 2007");
    printf("Output file = %s\n", a.out_path);
    if (a.flags & VS_F_NOISE) printf("SNR: %5.2f dB, ", 10.0 * log10((double)a.noise));
    printf("Fs=%ld Hz, Dur=%5.2f s, Fg=%d Hz, Amp = %d, DCflow=%5.2f\n", (long)a.fs, a.dur, (int)a.Fg, a.amp, a.DC);
    printf("Wait...");

    int dev = getenv("VS_DEVICE") ? atoi(getenv("VS_DEVICE")) : 0;
    vs_ctx *ctx = NULL;
    rc = vs_ctx_create(&ctx, &dev, 1, 0);
    if (rc) return die(NULL, rc, "no CUDA device");

    int16_t *pcm = (int16_t *)vs_host_alloc(n * sizeof *pcm);
    vs_period_rec *rec = (vs_period_rec *)calloc(max_periods ? max_periods : 1, sizeof *rec);
    if (!pcm || !rec) return die(ctx, VS_ENOMEM, "buffers");
    uint64_t rec_off[2] = {0, max_periods};
    uint32_t count = 0;
    vs_period_log log = {rec, rec_off, &count};
    const int want_lines = (a.flags & VS_F_NOISE) || ((a.flags & VS_F_SHIMMER) && a.shimmer != 0.0f);
    rc = vs_flowgen_batch(ctx, &p, 1, pcm, NULL, want_lines ? &log : NULL);
    if (rc) return die(ctx, rc, "vs_flowgen_batch");

    /* the reference prints these inside its period loop (:307, :409) */
    for (uint32_t k = 0; k < count; k++) {
        if ((a.flags & VS_F_SHIMMER) && a.shimmer != 0.0f) printf("%5.2f \n", rec[k].S);
        if (a.flags & VS_F_NOISE) printf("SNRdb = %5.2f\n", 10.0 * log10((double)(rec[k].x_pow / rec[k].w_pow)));
    }
    /* samples are little-endian int16 on every CUDA host */
    if (fwrite(pcm, sizeof *pcm, n, out) != n) { printf("Error while writing samples to %s\n", a.out_path); return 1; }
    fclose(out);
    vs_host_free(pcm);
    free(rec);
    vs_ctx_destroy(ctx);
    printf("done\n");
    return 0;
}
