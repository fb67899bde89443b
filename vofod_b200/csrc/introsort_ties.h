// libstdc++'s std::sort, restated step by step so that it can run in ONE device thread (or on the host) over a small array:
// pcl::EuclideanClusterExtraction orders its clusters with
//     std::sort(clusters.rbegin(), clusters.rend(), comparePointClusters);        // a.size() < b.size()   (extract_clusters.hpp, PCL 1.10)
// which is not stable: clusters of EQUAL size end up in whatever order introsort leaves.  That order decides which detection record
// carries which id (vofod_nodelet.cpp:840-843) whenever a scan has many clusters of the same size, so reproducing
// the reference there means reproducing the algorithm: introsort loop (median of three to the front, unguarded Hoare partition, recursion
// on the right part, 2*floor(log2 n) depth limit, heap sort below it) over blocks of > 16 elements, then the final insertion sort
// (guarded over the first 16, unguarded over the rest) — bits/stl_algo.h, bits/stl_heap.h (unchanged since GCC 4.x).
// Elements are moved as (key, payload) pairs; `less(a, b)` compares keys only.  The caller passes the sequence AS THE ALGORITHM SEES IT:
// for PCL's call through reverse iterators that is the cluster list reversed (and the result reversed back).
// Verified against std::sort itself by tests/cpp/introsort_vs_std_sort.cpp.  Not yet wired into the GPU ranking (next round).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define VF_HD __host__ __device__
#else
#define VF_HD
#endif

namespace introsort_ties
{
struct Item
{
  uint32_t key;      // cluster size
  uint32_t payload;  // which cluster
};
VF_HD inline bool less(const Item& a, const Item& b) { return a.key < b.key; }
VF_HD inline void swap_items(Item& a, Item& b)
{
  const Item t = a;
  a = b;
  b = t;
}

// ---- bits/stl_heap.h ----
VF_HD inline void push_heap_(Item* first, long hole, const long top, const Item value)
{
  long parent = (hole - 1) / 2;
  while (hole > top && less(first[parent], value))
  {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}
VF_HD inline void adjust_heap_(Item* first, long hole, const long len, const Item value)
{
  const long top = hole;
  long child = hole;
  while (child < (len - 1) / 2)
  {
    child = 2 * (child + 1);
    if (less(first[child], first[child - 1]))
      child--;
    first[hole] = first[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2)
  {
    child = 2 * (child + 1);
    first[hole] = first[child - 1];
    hole = child - 1;
  }
  push_heap_(first, hole, top, value);
}
#ifdef INTROSORT_TIES_COUNT_HEAP
static long g_heap_sorts = 0;  // (test aid: how often the depth limit was hit)
#endif
VF_HD inline void heap_sort_(Item* first, const long len)  // __partial_sort(first, last, last): make_heap + sort_heap
{
#ifdef INTROSORT_TIES_COUNT_HEAP
  g_heap_sorts++;
#endif
  if (len >= 2)
  {
    long parent = (len - 2) / 2;
    while (true)
    {
      const Item v = first[parent];
      adjust_heap_(first, parent, len, v);
      if (parent == 0)
        break;
      parent--;
    }
  }
  for (long last = len; last > 1; last--)  // __sort_heap: pop_heap(first, last)
  {
    const Item v = first[last - 1];
    first[last - 1] = first[0];
    adjust_heap_(first, 0, last - 1, v);
  }
}

// ---- bits/stl_algo.h ----
VF_HD inline void move_median_to_first_(Item* result, Item* a, Item* b, Item* c)
{
  if (less(*a, *b))
  {
    if (less(*b, *c)) swap_items(*result, *b);
    else if (less(*a, *c)) swap_items(*result, *c);
    else swap_items(*result, *a);
  } else if (less(*a, *c)) swap_items(*result, *a);
  else if (less(*b, *c)) swap_items(*result, *c);
  else swap_items(*result, *b);
}
VF_HD inline Item* unguarded_partition_(Item* first, Item* last, const Item* pivot)
{
  while (true)
  {
    while (less(*first, *pivot)) ++first;
    --last;
    while (less(*pivot, *last)) --last;
    if (!(first < last))
      return first;
    swap_items(*first, *last);
    ++first;
  }
}
VF_HD inline void unguarded_linear_insert_(Item* last)
{
  const Item val = *last;
  Item* next = last - 1;
  while (less(val, *next))
  {
    *last = *next;
    last = next;
    --next;
  }
  *last = val;
}
VF_HD inline void insertion_sort_(Item* first, Item* last)
{
  if (first == last)
    return;
  for (Item* i = first + 1; i != last; ++i)
  {
    if (less(*i, *first))
    {
      const Item val = *i;
      for (Item* p = i; p != first; --p)  // move_backward(first, i, i + 1)
        *p = *(p - 1);
      *first = val;
    } else
      unguarded_linear_insert_(i);
  }
}

// std::sort(first, first + n, less): the recursion of __introsort_loop (always on the RIGHT part, the left one continues the loop) becomes an
// explicit stack of (first, last, depth) — at most 2*log2(n) + 1 entries
VF_HD inline void sort(Item* first, const long n)
{
  if (n <= 0)
    return;
  long lg = 0;
  for (long v = n; v > 1; v >>= 1) lg++;
  struct Frame { long lo, hi, depth; };
  Frame stack[130];
  int sp = 0;
  stack[sp++] = Frame{0, n, 2 * lg};
  while (sp > 0)
  {
    Frame f = stack[--sp];
    while (f.hi - f.lo > 16)
    {
      if (f.depth == 0)
      {
        heap_sort_(first + f.lo, f.hi - f.lo);
        break;
      }
      f.depth--;
      Item* lo = first + f.lo;
      Item* hi = first + f.hi;
      Item* mid = lo + (hi - lo) / 2;
      move_median_to_first_(lo, lo + 1, mid, hi - 1);
      Item* cut = unguarded_partition_(lo + 1, hi, lo);
      // __introsort_loop(cut, last, depth): runs to completion BEFORE the loop goes on with [first, cut) — the order of the two does not
      // matter for the result (disjoint ranges), so the right part is simply queued
      stack[sp++] = Frame{(long)(cut - first), f.hi, f.depth};
      f.hi = (long)(cut - first);
    }
  }
  // __final_insertion_sort
  if (n > 16)
  {
    insertion_sort_(first, first + 16);
    for (Item* i = first + 16; i != first + n; ++i)
      unguarded_linear_insert_(i);
  } else
    insertion_sort_(first, first + n);
}

// PCL's call: sizes[] in seed order (cluster c = the c-th cluster found), order_out[r] = which cluster comes r-th.  scratch: n Items.
VF_HD inline void pcl_cluster_order(const uint32_t* sizes, const long n, Item* scratch, uint32_t* order_out)
{
  for (long i = 0; i < n; i++)
    scratch[i] = Item{sizes[n - 1 - i], (uint32_t)(n - 1 - i)};  // the sequence rbegin() .. rend()
  sort(scratch, n);
  for (long i = 0; i < n; i++)
    order_out[i] = scratch[n - 1 - i].payload;                   // back through the reverse iterators
}
}  // namespace introsort_ties
