// host_unpack.h — host-side half of the packed observation transfer of the host-buffer path (host_api.cu).
#pragma once
#include <cstddef>
#include <cstdint>

namespace dmfb {

// Persistent pool: n_threads - 1 workers plus the calling thread.
struct UnpackPool;
UnpackPool* unpack_pool_create(int n_threads);
void unpack_pool_destroy(UnpackPool* pool);

// Packed record of one agent's observation: ceil(cells / 2) bytes holding two 4-bit cells each (cell 2j in the low
// nibble of byte j), then the 2 direction bytes unchanged; records are `packed_stride` bytes apart.
inline size_t packed_record_bytes(int cells) { return ((size_t)(cells + 1) / 2 + 2 + 3) & ~(size_t)3; }

// Expands n_records packed records into int8 records of cells + 2 bytes (contiguous).  Blocks until done; the
// calling thread takes a share of the work.
void unpack_records(UnpackPool* pool, const uint8_t* packed, size_t packed_stride, int8_t* out, int cells,
                    size_t n_records);

}  // namespace dmfb
