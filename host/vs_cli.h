/* vs_cli.h -- argv handling of the two reference tools, shared by flowgen_shimmer, vowel and vs_batch.
 * Same flags, same value conversions, same acceptance ranges as the reference
 * (flowgen_shimmer.c:128-219,463-547; vowel_new.c:116-192). */
#ifndef VS_CLI_H
#define VS_CLI_H
#include <stdint.h>

typedef struct {
    /* one stream of vs_flow_params, in the units the library wants */
    float dur, jitter, shimmer, cq, K, Kvar, F0, DC, noise;
    int32_t amp, fs;
    uint8_t flags;
    /* shown in the banner only */
    float Fg;
    const char *out_path;
} vs_cli_flow;

typedef struct {
    const char *in_path, *out_path;
    int preset;              /* 'a','i','u','1'..'7' */
    float gain, pre;
    int has_noise;           /* -n given */
    float snr_linear;        /* pow(10, dB/10) */
} vs_cli_vowel;

/* both return 0, or -1 where the reference prints its usage text and exits.
 * need_out: the stand-alone tool insists on -o (flowgen_shimmer.c:219), the batch driver does not. */
int vs_cli_parse_flow(int argc, char **argv, int need_out, vs_cli_flow *out);
int vs_cli_parse_vowel(int argc, char **argv, vs_cli_vowel *out);
void vs_cli_flow_usage(void);
void vs_cli_vowel_usage(void);
/* srandom() seed: $VS_SEED if set, else time(NULL) like the reference (flowgen_shimmer.c:241) */
uint32_t vs_cli_seed(void);
#endif
