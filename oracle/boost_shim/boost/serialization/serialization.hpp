// Minimal stand-in so the unmodified MetaMaps sources compile without Boost.
// Test infrastructure only (oracle/_ref build); never part of the product.
#pragma once
#include <vector>
#include <map>
#include <unordered_map>
#include <string>
namespace boost { namespace serialization { class access; } }
