// Minimal stand-in for <boost/filesystem.hpp>: the reference's mj_model.h only declares path-typed globals.
#pragma once
#include <algorithm>
#include <filesystem>
#include <list>
#include <map>
#include <set>
#include <string>
#include <vector>  // the real header pulls these in transitively; the reference relies on that
namespace boost {
namespace filesystem {
using path = std::filesystem::path;
}  // namespace filesystem
}  // namespace boost
