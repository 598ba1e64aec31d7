#pragma once
#include "../../../../mm_refmath.h"
#include <stdexcept>
namespace boost { namespace math {
template <class RealType = double> struct chi_squared_distribution { RealType df_; chi_squared_distribution(RealType df) : df_(df) {} };
typedef chi_squared_distribution<double> chi_squared;
template <class T, class K> inline T cdf(const chi_squared_distribution<T>& d, const K& x) {
  if (d.df_ != 1) throw std::runtime_error("boost shim: chi_squared only for 1 df");
  return mmref::chi2_1df_cdf((double)x);
}
}}
