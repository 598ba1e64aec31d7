// Stand-in for boost/math/distributions/binomial.hpp (see oracle/mm_refmath.h).
#pragma once
#include "../../../../mm_refmath.h"
namespace boost { namespace math {
template <class RealType = double> struct binomial_distribution {
  RealType n_, p_;
  binomial_distribution(RealType n = 1, RealType p = 0.5) : n_(n), p_(p) {}
  RealType trials() const { return n_; }
  RealType success_fraction() const { return p_; }
};
typedef binomial_distribution<double> binomial;
template <class D, class R> struct complemented2 { const D& dist; R param; };
template <class D, class R> inline complemented2<D, R> complement(const D& d, const R& r) { return complemented2<D, R>{d, r}; }
template <class T, class K> inline T pdf(const binomial_distribution<T>& d, const K& k) { return mmref::binom_pmf((int)k, (int)d.n_, (double)d.p_); }
template <class T, class K> inline T cdf(const binomial_distribution<T>& d, const K& k) { return mmref::binom_cdf((int)std::floor((double)k), (int)d.n_, (double)d.p_); }
template <class T, class K> inline T cdf(const complemented2<binomial_distribution<T>, K>& c) { return mmref::binom_sf((int)std::floor((double)c.param), (int)c.dist.n_, (double)c.dist.p_); }
template <class T, class K> inline T quantile(const complemented2<binomial_distribution<T>, K>& c) { return mmref::binom_quantile_upper((int)c.dist.n_, (double)c.dist.p_, (double)c.param); }
}}
