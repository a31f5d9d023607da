// Shared host-side helpers of the C-ABI implementation: the error channel, and the vector type of the O(cells) host
// arrays.
#pragma once
#include <memory>
#include <string>
#include <utility>
#include <vector>

// Records `msg` as the calling thread's last error and returns `code` (a negative ma_status).
int ma_set_error(int code, const std::string &msg);

namespace ma {

// The O(cells) host arrays (layout, in-code mesh): a vector whose resize() leaves new elements uninitialised, so that
// the pages are first touched by the threads that fill them (the OpenMP loops of the builders, or big_assign) instead
// of by one thread value-initialising them: at 67 M cells that serial pass was half of the host builders' time.
template <class T>
struct NoInitAllocator : std::allocator<T> {
  template <class U>
  struct rebind {
    using other = NoInitAllocator<U>;
  };
  template <class U, class... A>
  void construct(U *p, A &&...a) {
    if constexpr (sizeof...(A) == 0)
      ::new ((void *)p) U;
    else
      ::new ((void *)p) U(std::forward<A>(a)...);
  }
};
template <class T>
using BigVec = std::vector<T, NoInitAllocator<T>>;
template <class T>
void big_assign(BigVec<T> &v, size_t n, T value) {
  v.clear();
  v.resize(n);
  T *p = v.data();
#pragma omp parallel for schedule(static)
  for (long i = 0; i < (long)n; ++i) p[i] = value;
}

}  // namespace ma
