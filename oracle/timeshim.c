/* TEST INFRASTRUCTURE ONLY.  LD_PRELOAD shim: the reference tools seed with srandom(time(NULL))
 * (flowgen_shimmer.c:241, vowel_new.c:234).  Returning $VS_SEED from time() makes the unmodified
 * binaries reproducible without touching their source. */
#include <stdlib.h>
#include <time.h>
time_t time(time_t *t)
{
    const char *s = getenv("VS_SEED");
    time_t v = s ? (time_t)atoll(s) : (time_t)1;
    if (t) *t = v;
    return v;
}
