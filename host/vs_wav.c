#include "vs_wav.h"
#include <stdlib.h>
#include <string.h>

static void put16(unsigned char *p, uint16_t v) { p[0] = (unsigned char)v; p[1] = (unsigned char)(v >> 8); }
static void put32(unsigned char *p, uint32_t v) { put16(p, (uint16_t)v); put16(p + 2, (uint16_t)(v >> 16)); }
static uint16_t get16(const unsigned char *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static uint32_t get32(const unsigned char *p) { return (uint32_t)get16(p) | ((uint32_t)get16(p + 2) << 16); }

int vs_wav_write_header(FILE *f, uint32_t sample_rate, uint32_t data_bytes)
{
    unsigned char h[44];
    memcpy(h, "RIFF", 4);
    put32(h + 4, data_bytes + 36);
    memcpy(h + 8, "WAVEfmt ", 8);
    put32(h + 16, 16);
    put16(h + 20, 1);                    /* PCM  */
    put16(h + 22, 1);                    /* mono */
    put32(h + 24, sample_rate);
    put32(h + 28, sample_rate * 2);      /* bytes per second */
    put16(h + 32, 2);                    /* block align */
    put16(h + 34, 16);                   /* bits per sample */
    memcpy(h + 36, "data", 4);
    put32(h + 40, data_bytes);
    return fwrite(h, 1, sizeof h, f) == sizeof h ? 0 : -1;
}

int vs_wav_read_header(FILE *f, vs_wav_info *info)
{
    unsigned char h[72];
    size_t got = fread(h, 1, sizeof h, f);
    if (got < 44 || memcmp(h, "RIFF", 4) != 0) return -1;
    memset(info, 0, sizeof *info);
    if (memcmp(h + 8, "WAVEfmt ", 8) == 0 && memcmp(h + 36, "data", 4) == 0) {
        /* canonical 44-byte header */
        info->format_tag = get16(h + 20);
        info->channels = get16(h + 22);
        info->sample_rate = get32(h + 24);
        info->bits_per_sample = get16(h + 34);
        info->data_bytes = get32(h + 40);
        info->data_offset = 44;
        return 0;
    }
    if (got == 72 && memcmp(h + 16, "WAVE", 4) == 0 && memcmp(h + 20, "fmt ", 4) == 0 && memcmp(h + 60, "data", 4) == 0) {
        /* the reference struct as laid out on LP64: 8-byte longs with padding (offsets measured with od) */
        info->format_tag = get16(h + 32);
        info->channels = get16(h + 34);
        info->sample_rate = get32(h + 40);
        info->bits_per_sample = get16(h + 58);
        info->data_bytes = get32(h + 64);
        info->data_offset = 72;
        return 0;
    }
    return -1;
}

int16_t *vs_wav_read_samples(FILE *f, const vs_wav_info *info, size_t *n_out)
{
    if (fseek(f, 0, SEEK_END) != 0) return NULL;
    long end = ftell(f);
    if (end < info->data_offset || fseek(f, info->data_offset, SEEK_SET) != 0) return NULL;
    size_t n = (size_t)(end - info->data_offset) / 2;
    int16_t *buf = (int16_t *)malloc((n ? n : 1) * sizeof *buf);
    if (!buf) return NULL;
    unsigned char *raw = (unsigned char *)buf;
    if (fread(raw, 2, n, f) != n) { free(buf); return NULL; }
    for (size_t i = 0; i < n; i++) buf[i] = (int16_t)get16(raw + 2 * i);     /* little-endian on any host */
    *n_out = n;
    return buf;
}
