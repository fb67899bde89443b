// TEST INFRASTRUCTURE ONLY — what the reference's src/pc_loader.cpp pulls in through <pcl/io/pcd_io.h>: the PCL console macros and
// the three boost string algorithms it calls (boost::trim, boost::split with is_any_of + token_compress_on).
#pragma once
#include <pcl/common/common.h>

#include <cerrno>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#define PCL_ERROR(...) do { } while (0)
#define PCL_INFO(...) do { } while (0)

namespace boost
{
// boost::trim: strips std::isspace characters (classic locale) from both ends
inline void trim(std::string& s)
{
  const char* ws = " \t\n\v\f\r";
  const std::size_t b = s.find_first_not_of(ws);
  if (b == std::string::npos) { s.clear(); return; }
  const std::size_t e = s.find_last_not_of(ws);
  s = s.substr(b, e - b + 1);
}
struct is_any_of { std::string set; explicit is_any_of(const char* s) : set(s) {} bool operator()(char c) const { return set.find(c) != std::string::npos; } };
enum token_compress_mode_type { token_compress_off, token_compress_on };
// boost::split: tokens between separator characters; with token_compress_on adjacent separators count as one.  Leading /
// trailing separators still yield an empty first / last token (the caller trims first, so none occur here).
template <class Pred>
inline void split(std::vector<std::string>& out, const std::string& in, Pred pred, token_compress_mode_type mode = token_compress_off)
{
  out.clear();
  std::string cur;
  std::size_t i = 0;
  while (i < in.size())
  {
    if (pred(in[i]))
    {
      out.push_back(cur);
      cur.clear();
      i++;
      if (mode == token_compress_on)
        while (i < in.size() && pred(in[i]))
          i++;
    } else
      cur.push_back(in[i++]);
  }
  out.push_back(cur);
}
}  // namespace boost
