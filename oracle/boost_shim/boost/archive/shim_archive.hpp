// Stand-in for Boost.Serialization archives. Only `metamaps index` /
// `mapAgainstIndex` touch archives; those sub-commands are outside the hot path,
// so every archive operation throws.
#pragma once
#include <istream>
#include <ostream>
#include <stdexcept>
namespace boost { namespace archive {
struct shim_archive_base {
  template <class T> shim_archive_base& operator&(T&)  { throw std::runtime_error("boost shim: serialization unavailable"); }
  template <class T> shim_archive_base& operator&(const T&)  { throw std::runtime_error("boost shim: serialization unavailable"); }
  template <class T> shim_archive_base& operator<<(const T&) { throw std::runtime_error("boost shim: serialization unavailable"); }
  template <class T> shim_archive_base& operator>>(T&) { throw std::runtime_error("boost shim: serialization unavailable"); }
};
struct text_oarchive   : shim_archive_base { explicit text_oarchive(std::ostream&) {} };
struct text_iarchive   : shim_archive_base { explicit text_iarchive(std::istream&) {} };
struct binary_oarchive : shim_archive_base { explicit binary_oarchive(std::ostream&) {} };
struct binary_iarchive : shim_archive_base { explicit binary_iarchive(std::istream&) {} };
}}
