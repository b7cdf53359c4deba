// Host-side copy between a caller's pageable buffer and the pinned staging of the host path (capi.cu: CopyPool).
// A plain memcpy of a 2 MiB slice writes through the cache: every destination line is first read for ownership, so a
// copy moves 3 bytes of DRAM traffic per byte.  Neither side is read again by this core (the DMA engine or the caller
// is next), so the destination is written with non-temporal stores: 2 bytes per byte.  Measured with 8-14 threads on the
// container's host (Xeon, 8 cores): 28 -> 37-42 GB/s.
#include <cstddef>
#include <cstdint>
#include <cstring>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace ccu {

#if defined(__x86_64__)
__attribute__((target("avx2"))) static void stream_copy_avx2(char* dst, const char* src, size_t n) {
  // head: up to the first 32-byte boundary of the destination
  const size_t head = (32 - (reinterpret_cast<uintptr_t>(dst) & 31)) & 31;
  if (head >= n) { std::memcpy(dst, src, n); return; }
  if (head) { std::memcpy(dst, src, head); dst += head; src += head; n -= head; }
  size_t i = 0;
  for (; i + 128 <= n; i += 128) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32));
    const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 64));
    const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 96));
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), a);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 32), b);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 64), c);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 96), d);
  }
  _mm_sfence();  // the stores are globally visible before the slice is reported done
  if (i < n) std::memcpy(dst + i, src + i, n - i);
}
#endif

// copy of one slice; `streaming` selects the non-temporal path where the CPU has it
void host_copy_slice(void* dst, const void* src, size_t bytes, bool streaming) {
#if defined(__x86_64__)
  static const bool has_avx2 = __builtin_cpu_supports("avx2");
  if (streaming && has_avx2 && bytes >= 4096) {
    stream_copy_avx2(static_cast<char*>(dst), static_cast<const char*>(src), bytes);
    return;
  }
#endif
  (void)streaming;
  std::memcpy(dst, src, bytes);
}

}  // namespace ccu
