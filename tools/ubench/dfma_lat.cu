// Dependent-issue latency of DFMA / IMAD / LOP3+IADD chains on the target GPU (one warp, clock64 deltas).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int n) {
    double a = out[0], b = out[1], acc = 0.0;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) { acc = fma(a, b, acc); acc = fma(b, a, acc); acc = fma(a, a, acc); acc = fma(b, b, acc); }
    long long t1 = clock64();
    int x = (int)out[2], y = (int)out[3], s = 0;
    long long t2 = clock64();
    for (int i = 0; i < n; i++) { s = s * x + y; s = s * y + x; s = s * x + y; s = s * y + x; }
    long long t3 = clock64();
    unsigned u = (unsigned)out[2], v = (unsigned)out[3];
    long long t4 = clock64();
    for (int i = 0; i < n; i++) { u = (u ^ v) + 0x9e3779b9u; u = __funnelshift_l(u, u, 7) + v; u = (u & v) + 0x7f4a7c15u; u = __funnelshift_l(u, u, 12) + v; }
    long long t5 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t3 - t2; cyc[2] = t5 - t4; out[4] = acc + s + u; }
}
int main() {
    double* d; long long* c; cudaMalloc(&d, 64); cudaMalloc(&c, 64);
    double h[8] = {1.0000001, 0.9999999, 3, 5, 0, 0, 0, 0}; cudaMemcpy(d, h, 64, cudaMemcpyHostToDevice);
    const int n = 4096;
    for (int w = 1; w <= 4; w *= 2) {
        k<<<1, 32 * w>>>(d, c, n); cudaDeviceSynchronize();
        long long hc[3]; cudaMemcpy(hc, c, 24, cudaMemcpyDeviceToHost);
        printf("warps/CTA %d: DFMA %.2f cyc/op, IMAD %.2f cyc/op, int op pair %.2f cyc/2ops\n", w, hc[0] / (4.0 * n), hc[1] / (4.0 * n), hc[2] / (4.0 * n));
    }
    return 0;
}
