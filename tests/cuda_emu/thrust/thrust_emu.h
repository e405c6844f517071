// CPU stand-in for the handful of thrust calls of pattern.cu (test infrastructure only).
#pragma once
#include <algorithm>
#include <numeric>
#include <vector>

#include "../cuda_runtime.h"

namespace thrust
{
  struct emu_policy
  {
    emu_policy on(cudaStream_t) const { return *this; }
  };
  namespace cuda
  {
    inline constexpr emu_policy par{};
  }
  template <class T>
  inline T *device_pointer_cast(T *p)
  {
    return p;
  }
  template <class T>
  inline void sequence(emu_policy, T *b, T *e)
  {
    std::iota(b, e, T(0));
  }
  template <class K, class V>
  inline void stable_sort_by_key(emu_policy, K *kb, K *ke, V *vb)
  {
    const size_t        n = size_t(ke - kb);
    std::vector<size_t> idx(n);
    std::iota(idx.begin(), idx.end(), size_t(0));
    std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return kb[a] < kb[b]; });
    std::vector<K> k2(n);
    std::vector<V> v2(n);
    for (size_t i = 0; i < n; ++i)
      {
        k2[i] = kb[idx[i]];
        v2[i] = vb[idx[i]];
      }
    std::copy(k2.begin(), k2.end(), kb);
    std::copy(v2.begin(), v2.end(), vb);
  }
  template <class In, class Out, class T>
  inline void exclusive_scan(emu_policy, In *b, In *e, Out *out, T init)
  {
    T acc = init;
    for (; b != e; ++b, ++out)
      {
        const T v = T(*b);
        *out      = Out(acc);
        acc += v;
      }
  }
} // namespace thrust
