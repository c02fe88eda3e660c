/* RAJA/RAJA.hpp -- SEQUENTIAL RAJA stand-in (test infrastructure, our own code).
 *
 * This container has no RAJA.  The reference's benchmarks/advection_reaction_3D is written
 * against RAJA; its serial configuration uses only Views, RangeSegments, a three-level
 * sequential kernel policy and forall.  This header implements exactly that subset with
 * plain loops, so that the UNMODIFIED reference sources compile by path into the CPU
 * program that produces the goldens and the CPU baseline for apps/advection_reaction_3D.
 * Elementwise kernels only: the iteration order has no effect on the results.
 */
#ifndef B200_SHIM_RAJA_HPP
#define B200_SHIM_RAJA_HPP

#include <cstddef>
#include <tuple>

#define RAJA_VERSION_MAJOR 2024
#define RAJA_VERSION_MINOR 7
#define RAJA_VERSION_PATCHLEVEL 0

namespace RAJA {

struct seq_exec {};
struct loop_exec {};

template <int N>
struct Layout {};

/* row-major view: the LAST index is contiguous (RAJA::Layout default permutation) */
template <typename T, typename L>
class View;

template <typename T, int N>
class View<T, Layout<N>>
{
public:
  template <typename... Ext>
  View(T* data, Ext... ext) : data_(data), ext_{static_cast<long>(ext)...}
  {
    static_assert(sizeof...(Ext) == N, "View: one extent per dimension");
  }

  template <typename... Idx>
  T& operator()(Idx... idx) const
  {
    static_assert(sizeof...(Idx) == N, "View: one index per dimension");
    const long ix[N] = {static_cast<long>(idx)...};
    long off         = 0;
    for (int d = 0; d < N; d++) off = off * ext_[d] + ix[d];
    return data_[off];
  }

private:
  T* data_;
  long ext_[N];
};

struct RangeSegment
{
  long b, e;
  RangeSegment(long begin, long end) : b(begin), e(end) {}
};

template <typename... T>
std::tuple<T...> make_tuple(T... t)
{
  return std::tuple<T...>(t...);
}

namespace statement {
template <int I, typename Exec, typename... Body>
struct For {};
template <int I>
struct Lambda {};
} // namespace statement

template <typename... S>
struct KernelPolicy {};

/* the one shape the benchmark uses: For<2, For<1, For<0, Lambda<0>>>> -- segment 2 is the
   outermost loop, segment 0 the innermost; the lambda takes (seg0, seg1, seg2) indices */
template <typename Policy, typename F>
void kernel(const std::tuple<RangeSegment, RangeSegment, RangeSegment>& segs, F&& f)
{
  const RangeSegment &s0 = std::get<0>(segs), &s1 = std::get<1>(segs), &s2 = std::get<2>(segs);
  for (long k = s2.b; k < s2.e; k++)
    for (long j = s1.b; j < s1.e; j++)
      for (long i = s0.b; i < s0.e; i++) f((int)i, (int)j, (int)k);
}

template <typename Exec, typename F>
void forall(const RangeSegment& r, F&& f)
{
  for (long i = r.b; i < r.e; i++) f((int)i);
}

} // namespace RAJA
#endif
