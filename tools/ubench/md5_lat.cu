// Dependent-issue latency of the MD5 step on sm_100a: pure chains of each instruction, then step variants.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/md5_lat tools/ubench/md5_lat.cu && /tmp/md5_lat
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP16(x) x x x x x x x x x x x x x x x x
#define REP256(x) REP16(REP16(x))

template <int V>
__global__ void chain(uint32_t* out, long long* cyc, uint32_t seed, int iters) {
    uint32_t a = seed + threadIdx.x, b = seed * 3u, c = seed * 5u, d = seed * 7u;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (V == 0) { REP256(asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));) }
        if (V == 1) { REP256(asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(b));) }
        if (V == 2) { REP256(asm volatile("{ .reg .u32 t; add.u32 t, %1, %2; add.u32 %0, %0, t; }" : "+r"(a) : "r"(b), "r"(c));) }   // may become IADD3 off-chain + IADD
        if (V == 3) { REP256(asm volatile("shf.l.wrap.b32 %0, %0, %0, 7;" : "+r"(a));) }
        if (V == 4) { REP256(asm volatile("{ .reg .u32 t; shf.l.wrap.b32 t, %0, %0, 7; add.u32 %0, t, %1; }" : "+r"(a) : "r"(b));) }   // LEA.HI
        if (V == 5) { REP256(asm volatile("mad.lo.u32 %0, %0, 1, %1;" : "+r"(a) : "r"(b));) }                                          // IMAD
        if (V == 6) { REP256(asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(a) : "r"(b), "r"(c));) }     // IADD3 on chain
    }
    long long t1 = clock64();
    out[threadIdx.x] = a + d;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// MD5 round-1-like steps, 256 per iteration; variants of how the step is formed
template <int V>
__global__ void steps(uint32_t* out, long long* cyc, const uint32_t* wsrc, int iters) {
    uint32_t a = wsrc[0] + threadIdx.x, b = wsrc[1], c = wsrc[2], d = wsrc[3];
    uint32_t w[16];
    for (int i = 0; i < 16; i++) w[i] = wsrc[4 + i];
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 64; i++) {
            uint32_t kw;
            asm volatile("add.u32 %0, %1, %2;" : "=r"(kw) : "r"(w[i & 15]), "r"(0xd76aa478u + i * 0x01010101u));
            uint32_t f = (i < 16) ? ((b & c) | (~b & d)) : (i < 32) ? ((d & b) | (~d & c)) : (i < 48) ? (b ^ c ^ d) : (c ^ (b | ~d));
            uint32_t t;
            if (V == 0) t = (a + kw) + f;                                        // compiler: IADD3
            if (V == 1) { uint32_t u; asm volatile("add.u32 %0, %1, %2;" : "=r"(u) : "r"(a), "r"(kw)); t = u + f; }   // 2-input add on the chain
            if (V == 2) { uint32_t u; asm volatile("add.u32 %0, %1, %2;" : "=r"(u) : "r"(a), "r"(kw)); asm volatile("mad.lo.u32 %0, %1, 1, %2;" : "=r"(t) : "r"(f), "r"(u)); }   // IMAD on the chain
            uint32_t r;
            const int s = 7 + (i & 3) * 5;
            if (V == 3) { t = (a + kw) + f; uint32_t rr; asm volatile("shf.l.wrap.b32 %0, %1, %1, %2;" : "=r"(rr) : "r"(t), "r"(s)); asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(rr), "r"(b)); }
            else r = b + __funnelshift_l(t, t, s);
            a = d; d = c; c = b; b = r;
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = a + b + c + d;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    uint32_t *out, *w; long long* cyc;
    cudaMalloc(&out, 4096); cudaMalloc(&w, 4096); cudaMalloc(&cyc, 64);
    cudaMemset(w, 0x5a, 4096);
    long long h;
    const char* names[] = {"LOP3 chain", "IADD chain", "IADD (+off-chain add)", "SHF chain", "rot+add (LEA.HI?)", "IMAD chain", "IADD3 chain"};
#define RUNC(V) chain<V><<<1, 32>>>(out, cyc, 1, 4); chain<V><<<1, 32>>>(out, cyc, 1, 64); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-26s %.2f cycles/op\n", names[V], (double)h / (64.0 * 256));
    RUNC(0) RUNC(1) RUNC(2) RUNC(3) RUNC(4) RUNC(5) RUNC(6)
    const char* sn[] = {"step: IADD3 + LEA.HI", "step: IADD + LEA.HI", "step: IMAD + LEA.HI", "step: IADD3 + SHF + IADD"};
#define RUNS(V, NT) steps<V><<<1, NT>>>(out, cyc, w, 4); steps<V><<<1, NT>>>(out, cyc, w, 256); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-26s %3d threads: %.2f cycles/step\n", sn[V], NT, (double)h / (256.0 * 64));
    RUNS(0, 32) RUNS(1, 32) RUNS(2, 32) RUNS(3, 32)
    RUNS(0, 64) RUNS(0, 128) RUNS(0, 256)
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
