#include <math.h>
#include <stdio.h>
#include <stdint.h>
/* exhaustive check: for every r in [0,2^31): r/d  ==  fma(fma(-q0,d,r), inv, q0) with q0 = r*inv, inv = fl(1/d) */
int main(void){
  const double ds[2] = {2147483647.0, 2147483647.0*10000.0};
  for (int t=0;t<2;t++){
    volatile double dv = ds[t]; const double d = dv; const double inv = 1.0/d;
    uint64_t bad=0, bad0=0;
    #pragma omp parallel for reduction(+:bad,bad0)
    for (int64_t r=0;r<2147483648LL;r++){
      double x=(double)r; double want = x/d; double q0 = x*inv; double rem = fma(-q0,d,x); double q = fma(rem,inv,q0);
      bad += (q!=want); bad0 += (q0!=want);
    }
    printf("d=%.1f mismatches with correction: %llu, plain multiply: %llu\n", d, (unsigned long long)bad, (unsigned long long)bad0);
  }
  return 0;
}
