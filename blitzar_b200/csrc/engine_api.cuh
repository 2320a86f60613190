// Type-erased per-curve entry points of the MSM engine. api.cu sees only this header, so the
// kernels of each curve are compiled exactly once, in that curve's own translation unit.
#pragma once
#include <cstdint>

#include "../../include/blitzar_b200.h"
#include "runtime.cuh"

namespace b200 {

struct MsmOptions {
  u32 window_bits = 0;  // 0 = choose from n
  u32 chunk1 = 32;      // chunk length of the first accumulation level
  u32 chunkn = 8;       // chunk length of the cascade levels
  u32 reduce_g1 = 8;    // bucket-reduction group size, first level (power of two)
  u32 reduce_gn = 8;    // bucket-reduction group size, later levels (power of two)
  u64 quad_threshold = 32768;  // launches with at most this many logical threads run 4 lanes each
};

struct EngineCtx {
  stream_t s;
  MsmOptions opt;
  const void* builtin;  // device-resident built-in ristretto generators g(0..num_builtin)
  uint64_t num_builtin;
};

struct Handle {
  unsigned curve_id;
  unsigned n;
  void* gens;  // device array of the curve's generator layout
};

template <class T> struct DevBuf {
  T* p = nullptr;
  stream_t s;
  DevBuf(size_t count, stream_t s_) : s(s_) { p = (T*)dev_alloc(count * sizeof(T), s); }
  ~DevBuf() { dev_free(p, s); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};

struct CurveVTable {
  unsigned curve_id, point_bytes, gen_bytes, abi_gen_bytes, abi_proj_bytes, abi_commit_bytes;
  void (*commit_device)(const EngineCtx&, void* out_commitments, void* out_partials, uint32_t num,
                        const sxt_sequence_descriptor* d, const void* generators_dev,
                        uint64_t offset_generators);
  void (*fixed_device)(const EngineCtx&, void* out_res, void* out_partials, const Handle* h,
                       int mode, unsigned element_num_bytes, const unsigned* bit_table,
                       const unsigned* lengths, unsigned num_outputs, unsigned n,
                       const uint8_t* scalars_dev);
  void (*ingest_projective)(const EngineCtx&, const void* raw_dev, void* gens, uint64_t n);
  void (*gens_to_projective)(const EngineCtx&, const void* gens, void* out_dev, uint64_t n);
  void (*store)(const EngineCtx&, const void* pts, void* out_dev, uint64_t count, bool commit);
  void (*sum_parts)(const EngineCtx&, const void* parts, uint32_t nparts, uint32_t count,
                    void* out_pts);
};
extern const CurveVTable kVTableEd25519, kVTableBls12381, kVTableBn254, kVTableGrumpkin;

// built-in ristretto generators g(first .. first+n) into the device generator layout
void launch_builtin_generators(const EngineCtx& ctx, void* gens, uint64_t first, uint64_t n);

}  // namespace b200
