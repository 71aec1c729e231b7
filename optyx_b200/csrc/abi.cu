// Error plumbing, device queries and the FP64 pipe micro-benchmarks of the C ABI.
#include <stdarg.h>
#include <string.h>
#include "b2_common.cuh"

namespace b2 {

static thread_local char g_err[512] = "";
static int g_sms[kMaxDevices] = {};   // per device
static thread_local const char* g_last_kernel = "";

void set_last_kernel(const char* name) { g_last_kernel = name; }

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return 1;
    }
    return 0;
}

// No CPU fallback: without a device every compute entry point fails here.
int current_device() {
    int dev = -1;
    return cudaGetDevice(&dev) == cudaSuccess ? dev : -1;
}

int num_sms() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { set_error("no CUDA device: %s", cudaGetErrorString(e)); return -1; }
    if (dev >= 0 && dev < kMaxDevices && g_sms[dev] > 0) return g_sms[dev];
    int sms = 0;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess || sms <= 0) {
        set_error("cannot query SM count: %s", cudaGetErrorString(e));
        return -1;
    }
    if (dev >= 0 && dev < kMaxDevices) g_sms[dev] = sms;
    return sms;
}

// ---- FP64 peak micro-benchmarks ---------------------------------------------
// Register-resident dependent chains, enough independent accumulators per
// thread to cover the pipe latency.  Used by bench.py to measure the
// denominators of the tensor-pipe roofline (MEASURED_PEAKS.json only has bf16).
__global__ void __launch_bounds__(256) dfma_peak_kernel(int iters, double* sink) {
    double a[16];
    const double x = 1.0000001, y = 1e-9 * (threadIdx.x + 1);
    #pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = y * (i + 1);
    for (int it = 0; it < iters; ++it) {
        #pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x, y);
    }
    double s = 0;
    #pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 123.456) sink[0] = s;
}

__global__ void __launch_bounds__(256) dmma884_peak_kernel(int iters, double* sink) {
    double c[8][2];
    #pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = 0; c[i][1] = 0; }
    double a = 1e-9 * (threadIdx.x + 1), b = 1.0000001;
    for (int it = 0; it < iters; ++it) {
        #pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile(
                "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    #pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) sink[0] = s;
}

__global__ void __launch_bounds__(256) dmma16816_peak_kernel(int iters, double* sink) {
    double c[4][4];
    #pragma unroll
    for (int i = 0; i < 4; ++i) { c[i][0] = 0; c[i][1] = 0; c[i][2] = 0; c[i][3] = 0; }
    double a[8], b[4];
    #pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1e-9 * (threadIdx.x + i + 1);
    #pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = 1.0000001 + i;
    for (int it = 0; it < iters; ++it) {
        #pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile(
                "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 "
                "{%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]),
                  "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
    double s = 0;
    #pragma unroll
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456) sink[0] = s;
}

}  // namespace b2

extern "C" int b2_abi_version(void) { return B2_ABI_VERSION; }
extern "C" const char* b2_last_error(void) { return b2::g_err; }
extern "C" int b2_device_sms(void) { return b2::num_sms(); }
extern "C" const char* b2_last_kernel(void) { return b2::g_last_kernel; }

extern "C" int b2_fp64_peak_kernel(int32_t kind, int32_t iters, double* sink_dev,
                                   double* flops_out, void* stream) {
    int sms = b2::num_sms();
    if (sms <= 0) return 1;
    cudaStream_t s = (cudaStream_t)stream;
    const int threads = 256, blocks = sms * 8;
    const double warps = (double)blocks * (threads / 32);
    if (kind == 0) {
        b2::dfma_peak_kernel<<<blocks, threads, 0, s>>>(iters, sink_dev);
        *flops_out = (double)blocks * threads * (double)iters * 16.0 * 2.0;
    } else if (kind == 1) {
        b2::dmma884_peak_kernel<<<blocks, threads, 0, s>>>(iters, sink_dev);
        *flops_out = warps * (double)iters * 8.0 * (2.0 * 8 * 8 * 4);
    } else if (kind == 2) {
        b2::dmma16816_peak_kernel<<<blocks, threads, 0, s>>>(iters, sink_dev);
        *flops_out = warps * (double)iters * 4.0 * (2.0 * 16 * 8 * 16);
    } else { b2::set_error("b2_fp64_peak_kernel: bad kind %d", kind); return 1; }
    return b2::check_launch("b2_fp64_peak_kernel");
}
