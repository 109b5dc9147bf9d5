/* vs_wav.h -- 16-bit mono PCM WAV I/O for the command-line tools.
 * Writer: canonical 44-byte RIFF header with fixed-width fields (what the reference's struct is on
 * ILP32, flowgen_shimmer.c:49-63,549-565).  Reader: sniffs the canonical layout and the 72-byte
 * layout the reference's `long`-typed struct has on LP64 (SURVEY.md discrepancy 4). */
#ifndef VS_WAV_H
#define VS_WAV_H
#include <stdint.h>
#include <stdio.h>

typedef struct {
    uint32_t sample_rate;
    uint16_t format_tag;       /* 1 = PCM */
    uint16_t channels;
    uint16_t bits_per_sample;
    uint32_t data_bytes;       /* as stated by the header (the reference trusts EOF, not this) */
    long     data_offset;      /* where the samples start */
} vs_wav_info;

/* data_bytes is what the header should claim; the reference writes (long)(dur*fs*2) (:555) */
int  vs_wav_write_header(FILE *f, uint32_t sample_rate, uint32_t data_bytes);
int  vs_wav_read_header(FILE *f, vs_wav_info *info);
/* reads every sample from data_offset to EOF (like the reference's fread loop, vowel_new.c:237) */
int16_t *vs_wav_read_samples(FILE *f, const vs_wav_info *info, size_t *n_out);
#endif
