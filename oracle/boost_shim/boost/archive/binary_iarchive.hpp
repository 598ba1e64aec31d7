#include "shim_archive.hpp"
