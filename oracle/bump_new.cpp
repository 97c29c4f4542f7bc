// oracle/bump_new.cpp -- TEST INFRASTRUCTURE: monotonic operator new for the PARITY build of the reference.
//
// The reference's DistributeOctTree sorts pair<int, ExtractorNode*> (ORBextractor.cpp:690), i.e. it breaks
// node-size ties by heap address, so with a general-purpose malloc its output depends on the allocation
// history of the process.  Linking this file makes small allocations strictly
// increasing in address and never reused, so "address order == creation order" and the unmodified reference
// becomes a pure function of its input.  That rule (the later-created node sorts higher) is the canonical
// oracle rule.  The timing build (liborbref.so) does NOT link this file.
#include <sys/mman.h>

#include <atomic>
#include <cstdlib>
#include <new>

namespace {
const size_t kSmall = 1024;              // list nodes (104 B) and small key vectors
const size_t kReserve = (size_t)32 << 30;  // virtual only (MAP_NORESERVE); touched pages are what is used
char* g_lo = nullptr;
std::atomic<size_t> g_used{0};
std::atomic<int> g_init{0};

void init_arena() {
  int expected = 0;
  if (g_init.compare_exchange_strong(expected, 1)) {
    void* p = mmap(nullptr, kReserve, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (p == MAP_FAILED) std::abort();
    g_lo = (char*)p;
    g_init.store(2);
  } else {
    while (g_init.load() != 2) {}
  }
}
}  // namespace

void* operator new(size_t n) {
  if (n <= kSmall) {
    if (g_init.load() != 2) init_arena();
    n = (n + 15) & ~(size_t)15;
    size_t off = g_used.fetch_add(n);
    if (off + n > kReserve) std::abort();
    return g_lo + off;
  }
  void* p = std::malloc(n);
  if (!p) throw std::bad_alloc();
  return p;
}
void* operator new[](size_t n) { return operator new(n); }
void operator delete(void* p) noexcept {
  if (!p) return;
  if (g_lo && (char*)p >= g_lo && (char*)p < g_lo + kReserve) return;  // arena memory is never reused
  std::free(p);
}
void operator delete[](void* p) noexcept { operator delete(p); }
void operator delete(void* p, size_t) noexcept { operator delete(p); }
void operator delete[](void* p, size_t) noexcept { operator delete(p); }
