/* vowel -- drop-in for the reference tool: WAV in, order-22 vocal-tract filter (vowel_new.c:252-296)
 * on the GPU through vs_vowel_filter_batch(), WAV out.  Reads both the canonical 44-byte header and
 * the 72-byte one the reference's 64-bit build writes; always writes the canonical one.
 *
 * -n (white noise added to the filtered output, vowel_new.c:302-324) is NOT part of the GPU hot path
 * (SURVEY.md 8f, row N1): this tool applies it on the host after the GPU filter, frame by frame, with
 * the C library's own srandom()/random() exactly as the reference does. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "voicesynth.h"
#include "vs_cli.h"
#include "vs_wav.h"

static int16_t round_half_down(double x)            /* vowel_new.c:413-427 */
{
    double dec = x - floor(x);
    if (dec > 0.5) x = x + 1;
    if (x > 32767) x = 32767; else if (x < -32767) x = -32767;
    return (int16_t)floor(x);
}

static void add_output_noise(int16_t *y, size_t n, float snr, uint32_t fs, uint32_t seed)
{
    const int ms1 = (int)(fs * 0.001 / 2.0) * 2;
    const size_t frame = (size_t)(50 * ms1);
    if (!frame) return;
    srandom(seed);
    for (size_t base = 0; base < n; base += frame) {
        const size_t ni = n - base < frame ? n - base : frame;
        float acc = 0.0f;
        for (size_t i = 0; i < ni; i++) acc += (float)y[base + i] * y[base + i];
        const float power = acc / (float)(short)ni;
        const float width = sqrt(12 * power / snr);
        for (size_t i = 0; i < ni; i++) {
            const float u = (1.0 * random()) / RAND_MAX;
            const float w = width * (u - 0.5);
            y[base + i] = round_half_down(1.0 * y[base + i] + 1.0 * w);
        }
    }
}

int main(int argc, char **argv)
{
    vs_cli_vowel a;
    if (vs_cli_parse_vowel(argc, argv, &a)) { vs_cli_vowel_usage(); return 0; }

    FILE *in = fopen(a.in_path, "rb");
    if (!in) { printf(".wav file not found\n"); return 1; }
    vs_wav_info wi;
    if (vs_wav_read_header(in, &wi)) { printf("%s: not a RIFF/WAVE file\n", a.in_path); return 1; }
    if (wi.format_tag != 1) { printf(".wav file is not PCM"); return 1; }
    if (wi.bits_per_sample != 16) printf(".wav file is not 16 bits per sample!");
    size_t n = 0;
    int16_t *flow = vs_wav_read_samples(in, &wi, &n);
    fclose(in);
    if (!flow) { printf("Error while reading %s\n", a.in_path); return 1; }

    FILE *out = fopen(a.out_path, "wb");
    if (!out) { printf("Error while creating file (%s)\n", a.out_path); return 1; }
    vs_wav_write_header(out, wi.sample_rate, wi.data_bytes);          /* the reference copies the input header */

    printf("Vocal-tract filter, vowel /%c/ -- libvoicesynth_cuda\n", a.preset);
    printf("pre_emphasis=%5.2f, gain=%5.2f, snr=%5.2f\nWait...", a.pre, a.gain, a.snr_linear);

    int dev = getenv("VS_DEVICE") ? atoi(getenv("VS_DEVICE")) : 0;
    vs_ctx *ctx = NULL;
    int rc = vs_ctx_create(&ctx, &dev, 1, 0);
    if (rc) { fprintf(stderr, "vowel: no CUDA device: %s\n", vs_strerror(rc)); return 1; }
    int16_t *pcm = (int16_t *)malloc((n ? n : 1) * sizeof *pcm);
    const uint8_t preset = (uint8_t)a.preset;
    const uint64_t ns = n;
    vs_filter_params f = {&preset, &a.gain, &a.pre};
    if (n) {
        rc = vs_vowel_filter_batch(ctx, flow, NULL, &ns, &f, 1, pcm, NULL, NULL);
        if (rc) { fprintf(stderr, "vowel: %s (%s)\n", vs_strerror(rc), vs_last_error(ctx)); return 1; }
    }
    if (a.has_noise) add_output_noise(pcm, n, a.snr_linear, wi.sample_rate, vs_cli_seed());
    if (fwrite(pcm, sizeof *pcm, n, out) != n) { printf("Error while writing %s\n", a.out_path); return 1; }
    fclose(out);
    free(pcm); free(flow);
    vs_ctx_destroy(ctx);
    printf("done\n");
    return 0;
}
