// Boost.Regex -> std::regex (ECMAScript grammar covers the two patterns MetaMaps uses).
#pragma once
#include <regex>
namespace boost {
using std::regex; using std::smatch; using std::regex_search; using std::regex_replace;
}
