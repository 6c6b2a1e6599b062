// extern "C" surface of libeditor_b200.so (see include/editor_b200.h).
#include "abi_internal.h"
#include <string.h>
#include <stdio.h>

namespace edb {
static thread_local char g_err[512] = "";
int edb_set_error(int code, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg ? msg : "");
    return code;
}
}  // namespace edb

extern "C" {

int edb_version(void) { return 100; }
const char* edb_last_error(void) { return edb::g_err; }

int edb_gemm_bf16(const EdbGemmDesc* d, void* stream) {
    if (d == nullptr) return edb::edb_set_error(EDB_ERR_SHAPE, "null descriptor");
    return edb::gemm_bf16(*d, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
