// Micro-benchmarks of the sm_100a pipes the pair kernel leans on: FP64 issue rates and the
// LSU wavefront cost of shared / global loads under partial broadcast.  Prints one line per
// test: name, cycles per warp-instruction per SM sub-partition (SMSP) and per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int OP>
__global__ void fp64_kernel(double *out, long long *cyc, double seed) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-6 + i;
    const double m = 1.0000001, k = 1e-7;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) a[i] = fma(a[i], m, k);
            if (OP == 1) a[i] = a[i] + k;
            if (OP == 2) a[i] = a[i] * m;
            if (OP == 3) a[i] = rint(a[i] * m);                 // DMUL + FRND
            if (OP == 4) { double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a[i])); a[i] = r + 1.5; }  // MUFU.RSQ64H + DADD
            if (OP == 5) a[i] = fmax(a[i] * m, k);              // DMUL + max
            if (OP == 6) a[i] = sqrt(a[i]) + 1.5;                // library sqrt + DADD
            if (OP == 7) a[i] = (a[i] + 6755399441055744.0) - 6755399441055744.0 + m;  // magic rint: 3 DADD
            if (OP == 8) a[i] = (double)__double2int_rn(a[i]) + m;  // F2I + I2F + DADD
            if (OP == 9) a[i] = (a[i] > 1.5) ? a[i] * m : a[i] + k; // DSETP + DMUL + DADD + SEL
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// Shared-memory loads: every lane loads WIDTH bytes; lanes are split into `groups` groups of
// consecutive lanes, each group reading one address; group addresses are `stride` bytes apart.
template <int WIDTH>
__global__ void lds_kernel(double *out, long long *cyc, int groups, int stride, int lane_stride) {
    extern __shared__ __align__(16) unsigned char sm[];
    for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) ((float *)sm)[i] = i * 1e-3f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int g = lane / (32 / groups);
    unsigned off = (g * stride + (lane % (32 / groups)) * lane_stride) & 32767u;
    off &= ~(unsigned)(WIDTH - 1);
    double acc0 = 0, acc1 = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const unsigned o = (off + u * 2048) & 32767u;
            if (WIDTH == 16) { double2 v = *(const double2 *)(sm + o); acc0 += v.x; acc1 += v.y; }
            if (WIDTH == 8) { double v = *(const double *)(sm + o); acc0 += v; }
            if (WIDTH == 4) { float v = *(const float *)(sm + o); acc0 += v; }
        }
        off = (off + (unsigned)(acc0 == 1.2345) * 16) & 32767u;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int WIDTH>
__global__ void ldg_kernel(const unsigned char *__restrict__ gm, double *out, long long *cyc, int groups, int stride, int lane_stride) {
    const int lane = threadIdx.x & 31;
    const int g = lane / (32 / groups);
    unsigned off = (g * stride + (lane % (32 / groups)) * lane_stride) & 32767u;
    off &= ~(unsigned)(WIDTH - 1);
    double acc0 = 0, acc1 = 0;
    // warm L1
    for (int i = threadIdx.x; i < 32768 / 8; i += blockDim.x) acc0 += __ldg((const double *)gm + i) * 1e-30;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const unsigned o = (off + u * 2048) & 32767u;
            if (WIDTH == 16) { double2 v = __ldg((const double2 *)(gm + o)); acc0 += v.x; acc1 += v.y; }
            if (WIDTH == 8) { double v = __ldg((const double *)(gm + o)); acc0 += v; }
        }
        off = (off + (unsigned)(acc0 == 1.2345) * 16) & 32767u;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void shfl_kernel(double *out, long long *cyc) {
    double a = threadIdx.x * 1e-3, b = a + 1;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            a += __shfl_down_sync(0xffffffffu, b, 1);
            b += __shfl_down_sync(0xffffffffu, a, 1);
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + b;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static double *d_out;
static long long *d_cyc;
static int n_sm;

static double Report(const char *name, int threads, double inst_per_iter) {
    long long h[1024];
    cudaDeviceSynchronize();
    cudaMemcpy(h, d_cyc, sizeof(long long) * n_sm, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < n_sm; ++i) avg += (double)h[i];
    avg /= n_sm;
    const int warps = threads / 32;
    // warp-instructions issued per SMSP = warps/4 * ITERS * inst_per_iter
    const double per_smsp = avg / ((warps / 4.0) * ITERS * inst_per_iter);
    printf("%-44s threads %4d  cycles/warp-inst/SMSP %7.3f   per SM %7.3f\n", name, threads, per_smsp, per_smsp / 4.0);
    return per_smsp;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    n_sm = prop.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", prop.name, n_sm, prop.clockRate);
    CK(cudaMalloc(&d_out, sizeof(double) * n_sm * 1024));
    CK(cudaMalloc(&d_cyc, sizeof(long long) * 1024));
    unsigned char *d_gm;
    CK(cudaMalloc(&d_gm, 65536));
    CK(cudaMemset(d_gm, 0, 65536));
    const int T = 1024;
    const char *fp_names[] = {"DFMA", "DADD", "DMUL", "DMUL+FRND(rint)", "MUFU.RSQ64H+DADD", "DMUL+DMNMX(fmax)", "sqrt()+DADD", "magic rint 2xDADD+DADD",
                              "F2I+I2F+DADD", "DSETP+DMUL+DADD+SEL"};
#define RUNFP(OP) fp64_kernel<OP><<<n_sm, T>>>(d_out, d_cyc, 1.0); CK(cudaGetLastError()); Report(fp_names[OP], T, 8.0);
    RUNFP(0) RUNFP(0) RUNFP(1) RUNFP(2) RUNFP(3) RUNFP(4) RUNFP(5) RUNFP(6) RUNFP(7) RUNFP(8) RUNFP(9)
    for (int t = 256; t <= 512; t *= 2) { fp64_kernel<0><<<n_sm, t>>>(d_out, d_cyc, 1.0); Report("DFMA (fewer warps)", t, 8.0); }
    shfl_kernel<<<n_sm, T>>>(d_out, d_cyc); Report("SHFL.DOWN 64-bit (2 SHFL.32)", T, 8.0);
    CK(cudaFuncSetAttribute(lds_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    const int groups_list[] = {1, 2, 4, 8, 16, 32};
    char name[128];
    for (int gi = 0; gi < 6; ++gi) {
        const int g = groups_list[gi];
        for (int stride : {128, 144}) {
            snprintf(name, sizeof name, "LDS.128 %2d distinct addr, stride %dB", g, stride);
            lds_kernel<16><<<n_sm, T, 32768>>>(d_out, d_cyc, g, stride, 0); CK(cudaGetLastError()); Report(name, T, 8.0);
        }
    }
    for (int gi = 0; gi < 6; ++gi) {
        const int g = groups_list[gi];
        snprintf(name, sizeof name, "LDS.64  %2d distinct addr, stride 136B", g);
        lds_kernel<8><<<n_sm, T, 32768>>>(d_out, d_cyc, g, 136, 0); CK(cudaGetLastError()); Report(name, T, 8.0);
    }
    lds_kernel<8><<<n_sm, T, 32768>>>(d_out, d_cyc, 1, 0, 8); Report("LDS.64  unit stride (32 x 8B)", T, 8.0);
    lds_kernel<16><<<n_sm, T, 32768>>>(d_out, d_cyc, 1, 0, 16); Report("LDS.128 unit stride (32 x 16B)", T, 8.0);
    lds_kernel<16><<<n_sm, T, 32768>>>(d_out, d_cyc, 1, 0, 32); Report("LDS.128 lane stride 32B (pp records)", T, 8.0);
    lds_kernel<16><<<n_sm, T, 32768>>>(d_out, d_cyc, 1, 0, 48); Report("LDS.128 lane stride 48B", T, 8.0);
    lds_kernel<4><<<n_sm, T, 32768>>>(d_out, d_cyc, 1, 0, 4); Report("LDS.32  unit stride", T, 8.0);
    for (int gi = 0; gi < 6; ++gi) {
        const int g = groups_list[gi];
        for (int stride : {128, 144}) {
            snprintf(name, sizeof name, "LDG.128 %2d distinct addr, stride %dB (L1 hit)", g, stride);
            ldg_kernel<16><<<n_sm, T>>>(d_gm, d_out, d_cyc, g, stride, 0); CK(cudaGetLastError()); Report(name, T, 8.0);
        }
    }
    ldg_kernel<8><<<n_sm, T>>>(d_gm, d_out, d_cyc, 1, 0, 8); Report("LDG.64  unit stride (L1 hit)", T, 8.0);
    ldg_kernel<16><<<n_sm, T>>>(d_gm, d_out, d_cyc, 1, 0, 16); Report("LDG.128 unit stride (L1 hit)", T, 8.0);
    return 0;
}
