// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Minimal stand-in for the subset of the Kokkos 2.x API that the miniAero
// reference sources use.  The reference (`/root/reference/kokkos/*.C`) does
// not vendor Kokkos and none is installed here, so `oracle/build_ref.sh`
// compiles the UNMODIFIED reference sources against this header into
// `oracle/_ref/`.  Nothing of the reference's arithmetic lives in Kokkos:
// it only supplies array storage, the flat parallel loop, atomics, a timer
// and a sort, which is all this file provides.
//
//   * View<D,...>      : LayoutRight storage, zero initialised, shared ownership
//   * parallel_for     : `#pragma omp parallel for schedule(static)` when built
//                        with -fopenmp, a plain loop otherwise
//   * deep_copy        : memcpy (mirrors alias their source) + a DUMP HOOK that
//                        writes the raw doubles/ints of every copied view to
//                        $MINIAERO_DUMP_DIR/<seq>_<label>.bin, which is how
//                        the tests obtain full-precision reference results
//                        without editing a line of the reference.
#pragma once
#include <algorithm>
#include <chrono>
#include <climits>
#include <limits>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <numeric>
#include <string>
#include <map>
#include <type_traits>
#include <typeinfo>
#include <vector>

#define KOKKOS_INLINE_FUNCTION inline
#define KOKKOS_FUNCTION
#define KOKKOS_LAMBDA [=]

namespace Kokkos {

struct LayoutRight {};
struct LayoutLeft {};
struct LayoutStride {};
enum MemoryTraitsFlags { Unmanaged = 1, RandomAccess = 2, Atomic = 4 };
template <unsigned T>
struct MemoryTraits {};

struct HostSpaceDevice {
  static void fence() {}
};
typedef HostSpaceDevice Serial;
typedef HostSpaceDevice OpenMP;
typedef HostSpaceDevice DefaultExecutionSpace;
typedef HostSpaceDevice DefaultHostExecutionSpace;

inline void initialize(int &, char **) {}
inline void initialize() {}
inline void finalize() {}

struct ALL {};

namespace standin {
// Data-type parsing: `T**[5][3]` is "array[5] of array[3] of T**": static extents
// are read outermost first with std::extent, dynamic rank is the pointer depth.
template <class T>
struct PtrDepth {
  typedef T value_type;
  enum { depth = 0 };
};
template <class T>
struct PtrDepth<T *> {
  typedef typename PtrDepth<T>::value_type value_type;
  enum { depth = PtrDepth<T>::depth + 1 };
};

struct Block {
  std::string label;
  std::vector<unsigned char> bytes;
};

inline int &dump_seq() {
  static int s = 0;
  return s;
}

inline void dump(const std::string &label, const void *p, size_t bytes, size_t elem, const size_t *dims,
                 int rank) {
  const char *dir = std::getenv("MINIAERO_DUMP_DIR");
  if (!dir || !p) return;
  // optional filter: only views whose label appears in the comma-separated $MINIAERO_DUMP_LABELS
  const char *only = std::getenv("MINIAERO_DUMP_LABELS");
  if (only && *only && (std::string(",") + only + ",").find("," + label + ",") == std::string::npos) return;
  char name[1024];
  std::snprintf(name, sizeof(name), "%s/%06d_%s.bin", dir, dump_seq()++, label.c_str());
  FILE *f = std::fopen(name, "wb");
  if (!f) return;
  // header: magic, element size, rank, dims[4] (all int64), then raw data
  int64_t hdr[7] = {0x4f52454132303031LL, (int64_t)elem, rank, 1, 1, 1, 1};
  for (int i = 0; i < rank && i < 4; ++i) hdr[3 + i] = (int64_t)dims[i];
  std::fwrite(hdr, sizeof(hdr), 1, f);
  std::fwrite(p, 1, bytes, f);
  std::fclose(f);
}
}  // namespace standin

template <class DataType, class... Props>
class View {
 public:
  typedef typename std::remove_all_extents<DataType>::type pointer_part;
  typedef typename standin::PtrDepth<pointer_part>::value_type value_type;
  typedef typename std::remove_const<value_type>::type non_const_value_type;
  enum {
    rank_dynamic = standin::PtrDepth<pointer_part>::depth,
    rank_static = std::rank<DataType>::value,
    Rank = rank_dynamic + rank_static
  };
  typedef View HostMirror;

  std::shared_ptr<standin::Block> block_;
  non_const_value_type *ptr_ = nullptr;
  size_t dim_[4] = {1, 1, 1, 1};
  size_t stride_[4] = {0, 0, 0, 0};

  View() { dim_[0] = 0; }

  explicit View(const std::string &label, size_t n0 = 0, size_t n1 = 0, size_t n2 = 0) {
    size_t dyn[3] = {n0, n1, n2};
    int r = 0;
    for (; r < rank_dynamic; ++r) dim_[r] = dyn[r];
    set_static(r);
    size_t total = 1;
    for (int i = 0; i < Rank; ++i) total *= dim_[i];
    size_t s = 1;
    for (int i = Rank - 1; i >= 0; --i) {
      stride_[i] = s;
      s *= dim_[i];
    }
    block_ = std::make_shared<standin::Block>();
    block_->label = label;
    block_->bytes.assign(total * sizeof(non_const_value_type) + 64, 0);  // zero initialised
    ptr_ = reinterpret_cast<non_const_value_type *>(block_->bytes.data());
  }

  // const / memory-trait / layout converting copy (shared ownership)
  template <class D2, class... P2>
  View(const View<D2, P2...> &o) : block_(o.block_), ptr_(const_cast<non_const_value_type *>(o.ptr_)) {
    static_assert(std::is_same<non_const_value_type, typename View<D2, P2...>::non_const_value_type>::value,
                  "value type mismatch");
    for (int i = 0; i < 4; ++i) {
      dim_[i] = o.dim_[i];
      stride_[i] = o.stride_[i];
    }
  }
  template <class D2, class... P2>
  View &operator=(const View<D2, P2...> &o) {
    block_ = o.block_;
    ptr_ = const_cast<non_const_value_type *>(o.ptr_);
    for (int i = 0; i < 4; ++i) {
      dim_[i] = o.dim_[i];
      stride_[i] = o.stride_[i];
    }
    return *this;
  }

  size_t dimension_0() const { return dim_[0]; }
  size_t dimension_1() const { return dim_[1]; }
  size_t extent(int i) const { return dim_[i]; }
  size_t size() const {
    size_t t = 1;
    for (int i = 0; i < Rank; ++i) t *= dim_[i];
    return t;
  }
  value_type *ptr_on_device() const { return ptr_; }
  value_type *data() const { return ptr_; }
  const std::string &label() const {
    static const std::string none("unallocated");
    return block_ ? block_->label : none;
  }

  value_type &operator()(size_t i0) const { return ptr_[i0 * stride_[0]]; }
  value_type &operator()(size_t i0, size_t i1) const { return ptr_[i0 * stride_[0] + i1 * stride_[1]]; }
  value_type &operator()(size_t i0, size_t i1, size_t i2) const {
    return ptr_[i0 * stride_[0] + i1 * stride_[1] + i2 * stride_[2]];
  }
  value_type &operator()(size_t i0, size_t i1, size_t i2, size_t i3) const {
    return ptr_[i0 * stride_[0] + i1 * stride_[1] + i2 * stride_[2] + i3 * stride_[3]];
  }
  value_type &operator[](size_t i0) const { return ptr_[i0 * stride_[0]]; }

 private:
  void set_static(int r) {
    if (rank_static > 0) dim_[r++] = std::extent<DataType, 0>::value;
    if (rank_static > 1) dim_[r++] = std::extent<DataType, 1>::value;
    if (rank_static > 2) dim_[r++] = std::extent<DataType, 2>::value;
  }
};

template <class V>
typename V::HostMirror create_mirror(const V &v) {
  return v;  // same memory space: the mirror aliases its source
}
template <class V>
typename V::HostMirror create_mirror_view(const V &v) {
  return v;
}

template <class VD, class VS>
void deep_copy(const VD &dst, const VS &src) {
  const size_t n = src.size();
  if (dst.ptr_ != src.ptr_ && n) {
    // every deep_copy in the reference is between identically shaped contiguous views
    std::memcpy((void *)dst.ptr_, (const void *)src.ptr_, n * sizeof(typename VS::non_const_value_type));
  }
  standin::dump(src.label(), src.ptr_, n * sizeof(typename VS::non_const_value_type),
                sizeof(typename VS::non_const_value_type), src.dim_, VS::Rank);
}

// subview(v, ALL(), j) of a rank-2 view -> strided rank-1 view
template <class D, class... P>
View<typename View<D, P...>::non_const_value_type *, LayoutStride, HostSpaceDevice> subview(const View<D, P...> &v,
                                                                                            ALL, size_t j) {
  View<typename View<D, P...>::non_const_value_type *, LayoutStride, HostSpaceDevice> s;
  s.block_ = v.block_;
  s.ptr_ = v.ptr_ + j * v.stride_[1];
  s.dim_[0] = v.dim_[0];
  s.stride_[0] = v.stride_[0];
  return s;
}

namespace standin {
// Optional launch log (bench.py's CPU baseline): with $MINIAERO_LAUNCH_LOG set, every parallel_for is
// timed; at exit the file receives one "functor <mangled type> <launches> <seconds>" line per functor
// type and one "step_end <seconds since first launch>" line per launch of TimeSolverExplicitRK4.h's
// `copy` functor (:134-153; launched once before the time loop, :335-336, then at the end of every
// time step, :488-489), which is how per-step times are read without editing the reference.
struct LaunchLog {
  bool on = false;
  std::string path;
  std::chrono::steady_clock::time_point t0;
  bool have_t0 = false;
  std::map<std::string, std::pair<long, double> > per_functor;
  std::vector<double> step_end;
  LaunchLog() {
    const char *p = std::getenv("MINIAERO_LAUNCH_LOG");
    if (p && *p) {
      on = true;
      path = p;
    }
  }
  ~LaunchLog() {
    if (!on) return;
    FILE *f = std::fopen(path.c_str(), "w");
    if (!f) return;
    for (auto &kv : per_functor)
      std::fprintf(f, "functor %s %ld %.9f\n", kv.first.c_str(), kv.second.first, kv.second.second);
    for (double t : step_end) std::fprintf(f, "step_end %.9f\n", t);
    std::fclose(f);
  }
  void record(const char *name, std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    if (!have_t0) {
      t0 = a;
      have_t0 = true;
    }
    auto &e = per_functor[name];
    e.first += 1;
    e.second += std::chrono::duration<double>(b - a).count();
    if (std::strstr(name, "4copyI")) step_end.push_back(std::chrono::duration<double>(b - t0).count());
  }
};
inline LaunchLog &launch_log() {
  static LaunchLog l;
  return l;
}
}  // namespace standin

template <class Functor>
void parallel_for(size_t n, const Functor &f) {
  const long nn = (long)n;
  standin::LaunchLog &log = standin::launch_log();
  std::chrono::steady_clock::time_point a;
  if (log.on) a = std::chrono::steady_clock::now();
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (long i = 0; i < nn; ++i) f((int)i);
  if (log.on) log.record(typeid(Functor).name(), a, std::chrono::steady_clock::now());
}

namespace Experimental {
template <class Scalar, class Device>
struct MinMax {
  struct value_type {
    Scalar min_val, max_val;
  };
  value_type *result_;
  explicit MinMax(value_type &r) : result_(&r) {}
};
}  // namespace Experimental

template <class Lambda, class Scalar, class Device>
void parallel_reduce(size_t n, const Lambda &f, Experimental::MinMax<Scalar, Device> red) {
  typename Experimental::MinMax<Scalar, Device>::value_type v;
  v.min_val = std::numeric_limits<Scalar>::max();
  v.max_val = std::numeric_limits<Scalar>::lowest();
  for (size_t i = 0; i < n; ++i) f((int)i, v);
  *red.result_ = v;
}

// ---- atomics (only used by the reference's -DATOMICS_FLUX build) ----------------------
inline void atomic_add(double *dst, double v) {
#ifdef _OPENMP
  uint64_t *p = reinterpret_cast<uint64_t *>(dst);
  uint64_t old_bits = __atomic_load_n(p, __ATOMIC_RELAXED), new_bits;
  do {
    double o;
    std::memcpy(&o, &old_bits, 8);
    double nv = o + v;
    std::memcpy(&new_bits, &nv, 8);
  } while (!__atomic_compare_exchange_n(p, &old_bits, new_bits, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
#else
  *dst += v;
#endif
}
inline double atomic_fetch_add(double *dst, double v) {
  double old = *dst;
  atomic_add(dst, v);
  return old;
}
// Kokkos 2.x semantics: returns the value found at *dst before the exchange attempt.
inline double atomic_compare_exchange(double *dst, double compare, double val) {
  uint64_t *p = reinterpret_cast<uint64_t *>(dst);
  uint64_t expected, desired;
  std::memcpy(&expected, &compare, 8);
  std::memcpy(&desired, &val, 8);
  __atomic_compare_exchange_n(p, &expected, desired, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  double found;
  std::memcpy(&found, &expected, 8);
  return found;
}

// ---- sort (Faces.h: BinSort on the left-cell id; only the iteration order depends on it) --
template <class KeyView>
struct BinOp1D {
  int nbins;
  double lo, hi;
  BinOp1D(int n, typename KeyView::non_const_value_type mn, typename KeyView::non_const_value_type mx)
      : nbins(n), lo((double)mn), hi((double)mx) {}
};

template <class KeyView, class Comp, class Device, class SizeType>
struct BinSort {
  KeyView keys;
  View<int *, Device> sort_order;
  BinSort(KeyView k, Comp, bool) : keys(k) {}
  void create_permute_vector() {
    const size_t n = keys.dimension_0();
    sort_order = View<int *, Device>("sort_order", n);
    std::vector<int> idx(n);
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return keys(a) < keys(b); });
    for (size_t i = 0; i < n; ++i) sort_order(i) = idx[i];
  }
};

namespace Impl {
struct Timer {
  std::chrono::steady_clock::time_point t0;
  Timer() : t0(std::chrono::steady_clock::now()) {}
  void reset() { t0 = std::chrono::steady_clock::now(); }
  double seconds() const {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
};
}  // namespace Impl

}  // namespace Kokkos
