// tmap.h — TMA descriptor construction helper (see tmap.cu)
#pragma once
#include "common.cuh"

namespace sl {

enum TmapDtype { TMAP_BF16 = 0, TMAP_F32 = 1 };

// dims[0] is the contiguous dimension; strides_bytes[i] is the byte stride of
// dims[i+1] (rank-1 entries).  Out-of-bounds box elements read as zero and are
// dropped on store — the kernels rely on this for TF "SAME" padding and ragged
// tile edges.
int make_tmap(CUtensorMap* out, TmapDtype dtype, int rank, const void* base, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128);

}  // namespace sl
