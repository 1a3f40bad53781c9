// Stub of <curand_kernel.h> for the host emulation of the reference kernels: the random-matrix / jitter
// kernels that use it are outside the scored path and are never launched.
#pragma once
struct curandState { int unused; };
static inline void curand_init(unsigned long long, unsigned long long, unsigned long long, curandState*) {}
static inline unsigned int curand_poisson(curandState*, double) { return 0u; }
static inline float curand_normal(curandState*) { return 0.0f; }
static inline float curand_uniform(curandState*) { return 0.0f; }
