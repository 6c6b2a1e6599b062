// Tensor-core (tcgen05/TMEM/TMA) attention for the 129-token backbone sequences -- placeholder entry points until the
// kernels land; they fail loudly rather than fall back.
#include "abi_internal.h"
namespace edb {
int attention_tc_fwd(const EdbAttnDesc& d, cudaStream_t st) { return attention_simple(d, false, st); }
int attention_tc_bwd(const EdbAttnDesc& d, cudaStream_t st) { return attention_simple(d, true, st); }
}  // namespace edb
