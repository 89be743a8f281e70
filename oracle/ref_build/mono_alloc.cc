// ORACLE — TEST INFRASTRUCTURE ONLY.
// Monotonic operator new for oracle/_ref/libref.so (linked -Bsymbolic, so only this library's own allocations come here).
// WHY: the reference's DistributeOctTree sorts vector<pair<int, ExtractorNode*>> (src/ORBextractor.cc:644-651): nodes
// with equally many keypoints are ordered by their HEAP ADDRESS, so the unchanged reference is reproducible only as far
// as the allocator is.  Inside a ref_* call every allocation is bump-allocated from one contiguous arena and nothing is
// reused, hence address order == creation order — the deterministic definition of that corner which the oracle and the
// CUDA kernel implement.  The arena is rewound at the start of the next call, when every allocation
// of the previous one is dead (pixel buffers of the pyramid levels, which outlive a call, use malloc — cvstub Mat).
#include <sys/mman.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <new>

namespace {
const size_t kArenaBytes = (size_t)8 << 30;  // virtual reservation; pages are touched only as far as a call needs
thread_local bool t_mono = false;
char* g_base = nullptr;
size_t g_pos = 0;

void* arena_alloc(size_t n, size_t align) {
  if (!g_base) {
    void* p = mmap(nullptr, kArenaBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (p == MAP_FAILED) { std::fprintf(stderr, "libref: arena mmap failed\n"); std::abort(); }
    g_base = (char*)p;
  }
  size_t a = (g_pos + align - 1) & ~(align - 1);
  if (a + n > kArenaBytes) { std::fprintf(stderr, "libref: arena exhausted\n"); std::abort(); }
  g_pos = a + n;
  return g_base + a;
}
inline bool in_arena(void* p) { return g_base && (char*)p >= g_base && (char*)p < g_base + kArenaBytes; }
inline void* alloc(size_t n, size_t align) {
  if (t_mono) return arena_alloc(n ? n : 1, align < 16 ? 16 : align);
  void* p = align <= 16 ? std::malloc(n ? n : 1) : std::aligned_alloc(align, (n + align - 1) / align * align);
  if (!p) throw std::bad_alloc();
  return p;
}
inline void dealloc(void* p) {
  if (p && !in_arena(p)) std::free(p);
}
}  // namespace

extern "C" void ref_mono_begin() {  // single-threaded use (the tests); rewinds the arena
  if (g_pos > ((size_t)256 << 20)) madvise(g_base, g_pos, MADV_DONTNEED);
  g_pos = 0;
  t_mono = true;
}
extern "C" void ref_mono_end() { t_mono = false; }

void* operator new(size_t n) { return alloc(n, 16); }
void* operator new[](size_t n) { return alloc(n, 16); }
void* operator new(size_t n, std::align_val_t a) { return alloc(n, (size_t)a); }
void* operator new[](size_t n, std::align_val_t a) { return alloc(n, (size_t)a); }
void operator delete(void* p) noexcept { dealloc(p); }
void operator delete[](void* p) noexcept { dealloc(p); }
void operator delete(void* p, size_t) noexcept { dealloc(p); }
void operator delete[](void* p, size_t) noexcept { dealloc(p); }
void operator delete(void* p, std::align_val_t) noexcept { dealloc(p); }
void operator delete[](void* p, std::align_val_t) noexcept { dealloc(p); }
void operator delete(void* p, size_t, std::align_val_t) noexcept { dealloc(p); }
void operator delete[](void* p, size_t, std::align_val_t) noexcept { dealloc(p); }
