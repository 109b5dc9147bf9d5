/* vowel -- drop-in for the reference tool: WAV in, order-22 vocal-tract filter (vowel_new.c:252-296)
 * on the GPU through vs_vowel_filter_batch(), WAV out.  Reads both the canonical 44-byte header and
 * the 72-byte one the reference's 64-bit build writes; always writes the canonical one.
 *
 * -n (white noise added to the filtered output, vowel_new.c:302-324) also runs on the
 * GPU: vs_vowel_noise_batch() on the filtered PCM, seeded like the reference with time() or $VS_SEED. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "voicesynth.h"
#include "vs_cli.h"
#include "vs_wav.h"

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

    /* the reference's banner (vowel_new.c:404-410) */
    printf(" \nMaurilio N. Vieira, 28 mar 97. \n");
    printf(" Cascade Formant Synthesiser\n");
    printf(" Formant frequencies/bandwithds from Rabiner & Schafer (1978),\n");
    printf(" Digital Processing of Speech Signals, Prentice Hall, pp. 74-77\n");
    printf("pre_emphasis=%5.2f, gain=%5.2f, snr=%5.2f\nWait...", a.pre, a.gain, a.snr_linear);

    int dev = getenv("VS_DEVICE") ? atoi(getenv("VS_DEVICE")) : 0;
    vs_ctx *ctx = NULL;
    int rc = vs_ctx_create(&ctx, &dev, 1, 0);
    if (rc) { fprintf(stderr, "vowel: no CUDA device: %s\n", vs_strerror(rc)); return 1; }
    int16_t *pcm = (int16_t *)malloc((n ? n : 1) * sizeof *pcm);
    const uint8_t preset = (uint8_t)a.preset;
    const uint64_t ns = n;
    vs_filter_params f = {&preset, &a.gain, &a.pre};
    /* -n: the output noise of a frame is scaled by the frame's power after quantisation (vowel_new.c:302-324), so a
     * sample that differs by one LSB would change every noise value of its frame: filter in the reference's own
     * operation order there (bit-exact); without -n the fast filter (+-1 LSB) is used */
    if (a.has_noise) vs_ctx_set_option(ctx, VS_OPT_EXACT_FILTER, 1.0);
    if (n) {
        rc = vs_vowel_filter_batch(ctx, flow, NULL, &ns, &f, 1, pcm, NULL, NULL);
        if (rc) { fprintf(stderr, "vowel: %s (%s)\n", vs_strerror(rc), vs_last_error(ctx)); return 1; }
    }
    if (a.has_noise && n) {
        const uint32_t seed = vs_cli_seed();
        const int32_t rate = (int32_t)wi.sample_rate;
        rc = vs_vowel_noise_batch(ctx, pcm, NULL, &ns, &a.snr_linear, &rate, &seed, 1);
        if (rc) { fprintf(stderr, "vowel: %s (%s)\n", vs_strerror(rc), vs_last_error(ctx)); return 1; }
    }
    if (fwrite(pcm, sizeof *pcm, n, out) != n) { printf("Error while writing %s\n", a.out_path); return 1; }
    fclose(out);
    free(pcm); free(flow);
    vs_ctx_destroy(ctx);
    printf("done\n");
    return 0;
}
