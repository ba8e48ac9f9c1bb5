// Host-side plumbing shared by the C-ABI translation units: error reporting, launch counting,
// per-device FFT tables.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "fft.cuh"
#include "salsa_b200.h"

namespace salsa {

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
int check_cuda(cudaError_t e, const char* what);
void count_launch(int n = 1);

#define SALSA_CUDA(expr)                                   \
    do {                                                   \
        int _rc = ::salsa::check_cuda((expr), #expr);      \
        if (_rc != SALSA_OK) return _rc;                   \
    } while (0)

// Twiddles and window of the 512-point transform in both precisions, resident on one device.
struct DeviceTables {
    FftTables<double> d;
    FftTables<float> f;
};

// Tables for the current device and the given window (n_fft doubles on the host).
int get_tables(const double* window_host, DeviceTables* out);

// Builds the effective window of a parameter block on the host (periodic Hann of win_len centred in
// n_fft, or the user's table).
void host_window(const salsa_params_t* p, double* out /* n_fft */);

int validate_params(const salsa_params_t* p, bool with_stft = true);

}  // namespace salsa
