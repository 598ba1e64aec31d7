#pragma once
#include "../../../../mm_refmath.h"
namespace boost { namespace math {
template <class RealType = double> struct poisson_distribution { RealType m_; poisson_distribution(RealType m = 1) : m_(m) {} };
typedef poisson_distribution<double> poisson;
template <class T, class K> inline T pdf(const poisson_distribution<T>& d, const K& k) { return mmref::poisson_pmf((int)k, (double)d.m_); }
}}
