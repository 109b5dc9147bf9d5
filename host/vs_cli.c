#include "vs_cli.h"
#include "voicesynth.h"
#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

uint32_t vs_cli_seed(void)
{
    const char *s = getenv("VS_SEED");
    return s && *s ? (uint32_t)strtoull(s, NULL, 10) : (uint32_t)time(NULL);
}

void vs_cli_flow_usage(void)
{
    puts("flowgen_shimmer (libvoicesynth_cuda) -- simulated glottal airflow after Fant (1979)\n"
         "usage: flowgen_shimmer -o file [-flag value ...]      (flags are case-insensitive)\n"
         "  -o x  output file (.wav, PCM, mono, 16 bits/sample)\n"
         "  -r x  sampling rate in Hz (22050)\n"
         "  -d x  duration in seconds, >= 0.5 (1.0)\n"
         "  -j x  jitter in percent (0)\n"
         "  -c x  closed quotient, 0..1 (0.55)\n"
         "  -f x  fundamental frequency F0 in Hz, >= 50 and < Fg (120)\n"
         "  -g x  glottal formant Fg in Hz, >= 50 (125)\n"
         "  -k x  speed of closure K, >= 0.5 (0.65)\n"
         "  -z x  variation of the speed of closure, 0..1 (0)\n"
         "  -s x  shimmer in percent, 0..100 (0)\n"
         "  -n x  cycle-to-cycle SNR in dB, 0..50: additive uniform noise in the closed phase\n"
         "  -a x  maximum amplitude, 0..32766 (12000)\n"
         "  -l x  DC flow as a fraction of the amplitude, 0..0.3 (0)\n"
         "environment: VS_SEED = srandom() seed (default: time), VS_DEVICE = CUDA device index");
}

void vs_cli_vowel_usage(void)
{
    puts("vowel (libvoicesynth_cuda) -- order-22 all-pole vocal tract filter\n"
         "usage: vowel -i in.wav -o out.wav -v x [-flag value ...]\n"
         "  -i x  input file  (.wav, PCM, mono, 16 bits/sample)\n"
         "  -o x  output file (.wav, PCM, mono, 16 bits/sample)\n"
         "  -v x  vowel preset: a i u (Rabiner & Schafer) or 1..7 (/a/ /e/ /e_/ /i/ /o_/ /o/ /u/)\n"
         "  -p x  pre-emphasis, 0..1 (1.0)\n"
         "  -g x  gain, >= 1 (10.0)\n"
         "  -n x  SNR in dB (> 0) of white noise added to the output");
}

int vs_cli_parse_flow(int argc, char **argv, int need_out, vs_cli_flow *o)
{
    /* value string per flag letter, NULL = not given */
    const char *val[26] = {0};
    memset(o, 0, sizeof *o);
    o->dur = 1.0f; o->cq = 0.55f; o->K = 0.65f; o->Fg = 125.0f; o->F0 = 120.0f; o->fs = 22050; o->amp = 12000;
    if (argc < 2) return -1;
    int i = 1;
    for (; i < argc && argv[i][0] == '-'; i += 2) {
        const int c = tolower((unsigned char)argv[i][1]);
        if (i + 1 >= argc) return -1;                         /* every flag takes a value */
        if (c < 'a' || c > 'z' || !strchr("ogfdcjknralzs", c)) return -1;
        if (c == 'n') o->DC = 0.25f;                          /* -n switches the DC flow on at parse time */
        val[c - 'a'] = argv[i + 1];
    }
    if (i < argc && argv[i][0] != 'i') return -1;             /* trailing garbage (the reference tolerates "i...") */
#define V(ch) val[(ch) - 'a']
    if (need_out && !V('o')) return -1;
    o->out_path = V('o');
    float f;
    /* the order below is the reference's: later checks read earlier results (Fg before F0, amp before DC) */
    if (V('d')) { f = (float)atof(V('d')); if (!(f >= 0.5)) return -1; o->dur = f; }
    if (V('j')) { f = (float)(atof(V('j')) / 100.0); if (!(f >= 0.0 && f <= 10.0)) return -1; o->jitter = f; }
    if (V('k')) { f = (float)atof(V('k')); if (!(f >= 0.50)) return -1; o->K = f; }
    if (V('c')) { f = (float)atof(V('c')); if (!(f >= 0.0 && f <= 1.0)) return -1; o->cq = f; }
    if (V('g')) { f = (float)atof(V('g')); if (!(f >= 50)) return -1; o->Fg = f; }
    if (V('f')) { f = (float)atof(V('f')); if (!(f >= 50 && f < o->Fg)) return -1; o->F0 = f; }
    if (V('n')) { f = (float)atof(V('n')); if (!(f >= 0.0 && f <= 50)) return -1; o->noise = (float)pow(10, (double)(f / 10)); }
    if (V('a')) { int v = atoi(V('a')); if (!(v >= 0 && v < 32767)) return -1; o->amp = v; }
    if (V('l')) { f = (float)atof(V('l')); if (!(f >= 0 && f <= 0.3)) return -1; o->DC = f * (float)o->amp; }
    if (V('z')) { f = (float)atof(V('z')); if (!(f >= 0 && f <= 1)) return -1; o->Kvar = f; }
    if (V('r')) { long l = atol(V('r')); if (l <= 0 || l > 2000000000L) return -1; o->fs = (int32_t)l; }
    if (V('s')) { f = (float)atof(V('s')); if (!(f >= 0 && f <= 100)) return -1; o->shimmer = f / 100; }
    o->flags = (uint8_t)((V('j') ? VS_F_JITTER : 0) | (V('s') ? VS_F_SHIMMER : 0) | (V('n') ? VS_F_NOISE : 0));
#undef V
    return 0;
}

int vs_cli_parse_vowel(int argc, char **argv, vs_cli_vowel *o)
{
    memset(o, 0, sizeof *o);
    o->gain = 10.0f; o->pre = 1.0f;
    if (argc < 2) return -1;
    int i = 1;
    for (; i < argc && argv[i][0] == '-'; i += 2) {
        if (i + 1 >= argc) return -1;
        const char *v = argv[i + 1];
        switch (tolower((unsigned char)argv[i][1])) {
        case 'p': o->pre = (float)atof(v); if (o->pre < 0.0 || o->pre > 1.0) return -1; break;
        case 'g': o->gain = (float)atof(v); if (o->gain < 1) return -1; break;
        case 'i': o->in_path = v; break;
        case 'o': o->out_path = v; break;
        case 'n': {
            float snr = (float)atof(v);
            if (snr <= 0) return -1;
            o->has_noise = 1;
            o->snr_linear = (float)pow(10, (double)(snr / 10));
            break;
        }
        case 'v':
            /* the reference lets upper-case A/I/U through and then filters nothing; we refuse them */
            if (!v[0] || !strchr("aiu1234567", v[0])) return -1;
            o->preset = v[0];
            break;
        default: return -1;
        }
    }
    if ((i < argc && argv[i][0] != 'i') || !o->in_path || !o->preset || !o->out_path) return -1;
    return 0;
}
