// Index introsort with the same decision sequence as libstdc++'s std::sort
// (median-of-three pivot to *first, unguarded Hoare partition, depth limit
// 2*floor(log2 n) with heapsort fallback, 16-element insertion-sort threshold),
// written from the published algorithm.  Needed because the reference orders
// detections and ground truths with an UNSTABLE std::sort on indices
// (det_matching.cc:44-51, 95-98): the order among equal keys decides which crowd
// GT a detection is assigned to, so bit-exact matching needs the same
// permutation, not just a valid sort.  Host+device so the host unit test can
// compare it against std::sort directly (tests/test_introsort.py).
#pragma once
#include <stdint.h>

#ifndef __CUDACC__
#ifndef __host__
#define __host__
#define __device__
#endif
#endif

namespace gn {

template <typename Less>
struct IntroSorter {
  int32_t* a;
  Less less;  // less(i, j): key[i] < key[j]

  __host__ __device__ void swp(int x, int y) { int32_t t = a[x]; a[x] = a[y]; a[y] = t; }

  __host__ __device__ void median_to_first(int result, int x, int y, int z) {
    if (less(a[x], a[y])) {
      if (less(a[y], a[z])) swp(result, y);
      else if (less(a[x], a[z])) swp(result, z);
      else swp(result, x);
    } else if (less(a[x], a[z])) swp(result, x);
    else if (less(a[y], a[z])) swp(result, z);
    else swp(result, y);
  }

  __host__ __device__ int partition(int first, int last, int pivot) {
    for (;;) {
      while (less(a[first], a[pivot])) ++first;
      --last;
      while (less(a[pivot], a[last])) --last;
      if (!(first < last)) return first;
      swp(first, last);
      ++first;
    }
  }

  // ---- heap fallback (std::__partial_sort(first, last, last)) ----------------
  __host__ __device__ void push_heap(int first, int hole, int top, int32_t value) {
    int parent = (hole - 1) / 2;
    while (hole > top && less(a[first + parent], value)) {
      a[first + hole] = a[first + parent];
      hole = parent;
      parent = (hole - 1) / 2;
    }
    a[first + hole] = value;
  }
  __host__ __device__ void adjust_heap(int first, int hole, int len, int32_t value) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
      child = 2 * (child + 1);
      if (less(a[first + child], a[first + (child - 1)])) --child;
      a[first + hole] = a[first + child];
      hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
      child = 2 * (child + 1);
      a[first + hole] = a[first + (child - 1)];
      hole = child - 1;
    }
    push_heap(first, hole, top, value);
  }
  __host__ __device__ void heap_sort(int first, int last) {
    const int len = last - first;
    if (len >= 2) {  // make_heap
      int parent = (len - 2) / 2;
      for (;;) {
        const int32_t v = a[first + parent];
        adjust_heap(first, parent, len, v);
        if (parent == 0) break;
        --parent;
      }
    }
    // heap_select's scan over [middle,last) is empty (middle == last); sort_heap:
    int end = last;
    while (end - first > 1) {
      --end;
      const int32_t v = a[end];
      a[end] = a[first];
      adjust_heap(first, 0, end - first, v);
    }
  }

  __host__ __device__ void linear_insert(int last) {
    const int32_t v = a[last];
    int next = last - 1;
    while (less(v, a[next])) {
      a[last] = a[next];
      last = next;
      --next;
    }
    a[last] = v;
  }
  __host__ __device__ void insertion_sort(int first, int last) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
      if (less(a[i], a[first])) {
        const int32_t v = a[i];
        for (int j = i; j > first; --j) a[j] = a[j - 1];
        a[first] = v;
      } else {
        linear_insert(i);
      }
    }
  }

  __host__ __device__ void sort(int n) {
    if (n <= 1) return;
    // introsort loop with an explicit stack (device code: no recursion)
    int depth0 = 0;
    for (int m = n; m > 1; m >>= 1) ++depth0;
    depth0 *= 2;
    int stack_first[64], stack_last[64], stack_depth[64];
    int sp = 0;
    stack_first[0] = 0; stack_last[0] = n; stack_depth[0] = depth0; sp = 1;
    while (sp > 0) {
      --sp;
      int first = stack_first[sp], last = stack_last[sp], depth = stack_depth[sp];
      // libstdc++ recurses on the RIGHT part and loops on the left; the two
      // parts are disjoint, so processing order does not change the result.
      while (last - first > 16) {
        if (depth == 0) {
          heap_sort(first, last);
          break;
        }
        --depth;
        const int mid = first + (last - first) / 2;
        median_to_first(first, first + 1, mid, last - 1);
        const int cut = partition(first + 1, last, first);
        if (sp < 64) {
          stack_first[sp] = cut; stack_last[sp] = last; stack_depth[sp] = depth; ++sp;
        }
        last = cut;
      }
    }
    // final insertion sort
    if (n > 16) {
      insertion_sort(0, 16);
      for (int i = 16; i < n; ++i) linear_insert(i);
    } else {
      insertion_sort(0, n);
    }
  }
};

}  // namespace gn
