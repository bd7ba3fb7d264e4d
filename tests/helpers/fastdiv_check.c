/* CPU mirror of div_tab() in scema_b200/csrc/resample.cu (K1): a/b for a table divisor b with
 * rb = RN(1/b) as q = a*rb followed by two FMA corrections, compared bit for bit with the IEEE
 * quotient. Divisors are exactly the ones K1 meets: the knot spacings hd_i = x_{i+1} - x_i and the
 * eliminated diagonal di_i of the natural-spline system (spline.h:302-313, :195-219) for every
 * history length 3..Lmax. Numerators: random mantissas over 2^-850..2^850 and adversarial values
 * next to exact products q*b and to rounding midpoints. Build with -mfma -ffp-contract=off.
 * usage: fastdiv_check Lmax reps  -> prints totals, exit status 1 on any mismatch of the 2-step form. */
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
static uint64_t s = 0x9E3779B97F4A7C15ull;
static uint64_t rnd(void){ s += 0x9E3779B97F4A7C15ull; uint64_t z=s; z=(z^(z>>30))*0xBF58476D1CE4E5B9ull; z=(z^(z>>27))*0x94D049BB133111EBull; return z^(z>>31);}
static inline double fastdiv(double a, double b, double rb){
    double q = a*rb; double r = fma(-b,q,a); q = fma(r,rb,q); r = fma(-b,q,a); return fma(r,rb,q);
}
static inline double fastdiv1(double a, double b, double rb){
    double q = a*rb; double r = fma(-b,q,a); return fma(r,rb,q);
}
int main(int argc,char**argv){
    long bad=0, bad1=0, total=0;
    int Lmax = argc>1?atoi(argv[1]):2048;
    int reps = argc>2?atoi(argv[2]):200;
    for(int L=3; L<=Lmax; L++){
        // divisors: hd[i] = x[i+1]-x[i], di[] from the elimination
        double *x=malloc(sizeof(double)*L), *lo=malloc(sizeof(double)*L), *di=malloc(sizeof(double)*L), *up=malloc(sizeof(double)*L);
        for(int i=0;i<L;i++) x[i]=(double)i/(double)(L-1);
        const double third=1.0/3.0, twothird=2.0/3.0;
        for(int i=1;i<L-1;i++){ lo[i]=third*(x[i]-x[i-1]); di[i]=twothird*(x[i+1]-x[i-1]); up[i]=third*(x[i+1]-x[i]); }
        di[0]=2; up[0]=0; lo[0]=0; di[L-1]=2; lo[L-1]=0; up[L-1]=0;
        for(int i=0;i<L;i++){ double sd=1.0/di[i]; if(i>0) lo[i]*=sd; if(i<L-1) up[i]*=sd; di[i]=1.0; }
        for(int k=0;k<L-1;k++){ double xx=-lo[k+1]/di[k]; lo[k+1]=-xx; di[k+1]=di[k+1]+xx*up[k]; }
        for(int i=0;i<L;i++){
            for(int which=0; which<2; which++){
                double b = which? di[i] : (i<L-1? x[i+1]-x[i] : 1.0);
                double rb = 1.0/b;
                for(int r=0;r<reps;r++){
                    uint64_t m = rnd();
                    // random mantissa, exponent in a wide range
                    int e = (int)(rnd()%1700) - 850;
                    double a = ldexp(1.0 + (double)(m>>12)*0x1p-52, e); if(m&1) a=-a;
                    if (r & 1) {  // adversarial: numerators next to exact products q*b and midpoints
                        double q = ldexp(1.0 + (double)(rnd()>>12)*0x1p-52, e);
                        double half = (r & 2) ? ldexp(1.0, e-53) : 0.0;
                        a = (q + half) * b;   // rounded product
                        int k = (int)(rnd()%5) - 2;
                        for (int t=0;t<abs(k);t++) a = nextafter(a, k>0? INFINITY : -INFINITY);
                        if (m&1) a=-a;
                    }
                    double want=a/b, got=fastdiv(a,b,rb), g1=fastdiv1(a,b,rb);
                    total++;
                    if(memcmp(&want,&got,8)) { if(bad<5) printf("BAD a=%a b=%a want=%a got=%a\n",a,b,want,got); bad++; }
                    if(memcmp(&want,&g1,8)) bad1++;
                }
            }
        }
        free(x);free(lo);free(di);free(up);
    }
    printf("total=%ld bad(2-step)=%ld bad(1-step)=%ld\n", total, bad, bad1);
    return bad!=0;
}
