// vofod_b200/csrc/introsort_ties.h against libstdc++'s std::sort itself, called exactly as PCL calls it (reverse iterators, compare by size):
// random arrays with few distinct sizes (many ties), sorted / reversed / organ-pipe / all-equal inputs, and inputs that drive introsort into its
// heap-sort fallback.  Exit code 0 = identical order in every case.
#include <algorithm>
#include <cstdio>
#include <random>
#include <vector>

#define INTROSORT_TIES_COUNT_HEAP
#include "../../vofod_b200/csrc/introsort_ties.h"

struct Cl
{
  std::vector<int> indices;  // like pcl::PointIndices: only the size is compared
  unsigned id;
};

static bool check(const std::vector<uint32_t>& sizes, const char* what)
{
  const long n = (long)sizes.size();
  std::vector<Cl> cl((size_t)n);
  for (long i = 0; i < n; i++)
  {
    cl[(size_t)i].indices.resize(sizes[(size_t)i]);
    cl[(size_t)i].id = (unsigned)i;
  }
  std::sort(cl.rbegin(), cl.rend(), [](const Cl& a, const Cl& b) { return a.indices.size() < b.indices.size(); });
  std::vector<introsort_ties::Item> scratch((size_t)n);
  std::vector<uint32_t> order((size_t)n);
  introsort_ties::pcl_cluster_order(sizes.data(), n, scratch.data(), order.data());
  for (long i = 0; i < n; i++)
    if (order[(size_t)i] != cl[(size_t)i].id)
    {
      std::printf("MISMATCH %s n=%ld at rank %ld: ours %u, std::sort %u\n", what, n, i, order[(size_t)i], cl[(size_t)i].id);
      return false;
    }
  return true;
}

int main()
{
  std::mt19937 rng(12345);
  long cases = 0;
  bool ok = true;
  for (int rep = 0; rep < 3000 && ok; rep++)
  {
    const long n = 1 + (long)(rng() % (rep < 2000 ? 400 : 6000));
    const uint32_t distinct = 1 + rng() % (rep % 3 == 0 ? 3 : (rep % 3 == 1 ? 12 : 200));
    std::vector<uint32_t> s((size_t)n);
    for (auto& v : s) v = 1 + rng() % distinct;
    ok = ok && check(s, "random");
    cases++;
  }
  for (long n : {0L, 1L, 2L, 15L, 16L, 17L, 18L, 31L, 32L, 33L, 100L, 1000L, 4097L})
  {
    std::vector<uint32_t> s((size_t)n);
    for (long i = 0; i < n; i++) s[(size_t)i] = (uint32_t)(i + 1);
    ok = ok && check(s, "ascending");
    std::reverse(s.begin(), s.end());
    ok = ok && check(s, "descending");
    for (long i = 0; i < n; i++) s[(size_t)i] = (uint32_t)(i < n / 2 ? i : n - i);
    ok = ok && check(s, "organ pipe");
    for (long i = 0; i < n; i++) s[(size_t)i] = 7;
    ok = ok && check(s, "all equal");
    for (long i = 0; i < n; i++) s[(size_t)i] = (uint32_t)(i % 2 ? 3 : 9);
    ok = ok && check(s, "alternating");
    cases += 5;
  }
  // median-of-three killer (Musser): drives the depth limit to 0 => heap sort inside std::sort
  for (long n : {64L, 200L, 1000L, 5000L})
  {
    std::vector<uint32_t> s((size_t)n);
    const long k = n / 2;
    for (long i = 1; i <= k; i++)
    {
      s[(size_t)(i - 1)] = (uint32_t)(i % 2 ? i : k + i - 1);
      s[(size_t)(k + i - 1)] = (uint32_t)(2 * i);
    }
    ok = ok && check(s, "median-of-3 killer");
    std::reverse(s.begin(), s.end());
    ok = ok && check(s, "median-of-3 killer, reversed");
    cases += 2;
  }
  std::printf("%s: %ld cases, %ld of them through the heap-sort fallback\n", ok ? "introsort_ties == std::sort (PCL's call)" : "FAILED", cases,
              introsort_ties::g_heap_sorts);
  return ok && introsort_ties::g_heap_sorts > 0 ? 0 : 1;
}
