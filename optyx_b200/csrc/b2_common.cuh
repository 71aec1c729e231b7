// Shared helpers for the optyx_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/optyx_b200.h"

namespace b2 {

template <typename T> struct cplx;
template <> struct cplx<double> { using type = double2; };
template <> struct cplx<float> { using type = float2; };

template <typename T>
__device__ __forceinline__ typename cplx<T>::type make_c(T re, T im) {
    typename cplx<T>::type r; r.x = re; r.y = im; return r;
}

// acc += a * b   (4 FMAs)
template <typename C>
__device__ __forceinline__ void cfma(C& acc, const C a, const C b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}

template <typename C>
__device__ __forceinline__ C cmul(const C a, const C b) {
    C r;
    r.x = a.x * b.x - a.y * b.y;
    r.y = a.x * b.y + a.y * b.x;
    return r;
}

// "done once PER DEVICE" flags (kernel function attributes are per device, and one
// process may drive several devices)
constexpr int kMaxDevices = 64;
struct once_flags { bool done[kMaxDevices] = {}; };
int current_device();                       // -1 without a device
inline bool is_done(const once_flags& f) {
    const int d = current_device();
    return d >= 0 && d < kMaxDevices && f.done[d];
}
inline void mark_done(once_flags& f) {
    const int d = current_device();
    if (d >= 0 && d < kMaxDevices) f.done[d] = true;
}

void set_error(const char* fmt, ...);
int check_launch(const char* what);
int num_sms();
void set_last_kernel(const char* name);

}  // namespace b2
