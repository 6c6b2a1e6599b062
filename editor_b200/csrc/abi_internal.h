// Internal declarations shared by the .cu translation units behind include/editor_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include "../../include/editor_b200.h"

namespace edb {

int edb_set_error(int code, const char* msg);

#define EDB_CHECK_LAUNCH()                                                              \
    do {                                                                                \
        cudaError_t e__ = cudaGetLastError();                                           \
        if (e__ != cudaSuccess) return edb::edb_set_error(EDB_ERR_CUDA, cudaGetErrorString(e__)); \
    } while (0)

#define EDB_TRY(expr)                  \
    do {                               \
        int rc__ = (expr);             \
        if (rc__ != EDB_OK) return rc__; \
    } while (0)

int num_sms();
int make_tmap_bf16(CUtensorMap* map, const void* base, long long inner, long long outer, long long ld, int box_rows);
int gemm_bf16(const EdbGemmDesc& g, cudaStream_t stream);

}  // namespace edb
