// Micro-probe (not part of the library): dependent-chain latency of DADD / DFMA / the Neumaier step on sm_100.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(double *out, long long *cyc, int n, double x0)
{
    double s = 0.0, c = 0.0, x = x0;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) s = __dadd_rn(s, x);
    long long t1 = clock64();
    double f = 1.0;
    for (int i = 0; i < n; ++i) f = __fma_rn(f, x, x);
    long long t2 = clock64();
    double s2 = 0.0;
    for (int i = 0; i < n; ++i) {
        double xx = x * (double)(i & 7);
        double t = __dadd_rn(s2, xx);
        bool big = fabs(s2) >= fabs(xx);
        double hi = big ? s2 : xx, lo = big ? xx : s2;
        c = __dadd_rn(c, __dadd_rn(__dadd_rn(hi, -t), lo));
        s2 = t;
    }
    long long t3 = clock64();
    float g = 0.f, y = (float)x0;
    for (int i = 0; i < n; ++i) g = __fadd_rn(g, y);
    long long t4 = clock64();
    out[0] = s + f + s2 + c + g;
    cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3;
}
int main()
{
    double *o; long long *c; cudaMalloc(&o, 8); cudaMalloc(&c, 32);
    int n = 20000;
    for (int it = 0; it < 2; ++it) probe<<<1, 32>>>(o, c, n, 1.0000001);
    long long h[4]; cudaMemcpy(h, c, 32, cudaMemcpyDeviceToHost);
    printf("DADD chain %.1f cyc/op, DFMA chain %.1f, Neumaier step %.1f, FADD chain %.1f\n", h[0] / (double)n, h[1] / (double)n, h[2] / (double)n, h[3] / (double)n);
    return 0;
}
