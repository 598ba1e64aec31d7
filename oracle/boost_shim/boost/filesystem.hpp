// Stand-in for the three Boost.Filesystem calls MetaMaps makes.
#pragma once
#include <string>
#include <sstream>
#include <set>
#include <map>
#include <cstdio>
#include <sys/stat.h>
namespace boost { namespace filesystem {
struct wpath { std::string p; wpath(const std::string& s) : p(s) {} };
inline bool remove(const wpath& w) { return std::remove(w.p.c_str()) == 0; }
inline bool exists(const std::string& s) { struct stat st; return ::stat(s.c_str(), &st) == 0; }
}}
