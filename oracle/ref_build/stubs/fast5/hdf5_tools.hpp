// Stub of the tiny part of hdf5_tools::File that src/read_buffer.cpp touches
// (get_attr_map / read).  It lets oracle/_ref be built from the reference's own
// mapper sources with no HDF5 library; the HDF5 ReadBuffer constructor compiles but is
// never called by the shim (signals are passed in as float arrays).
// Test infrastructure only -- not part of the product.
#pragma once
#include <map>
#include <string>
#include <vector>
#include <array>
#include <deque>
#include <iostream>
#include <sstream>
#include <fstream>
#include <cassert>
#include <exception>
#include <functional>
#include <algorithm>
#include <cstring>
#include <mutex>
#include <thread>
#include <chrono>
#include <limits>
#include <memory>
#include <queue>
#include <set>
namespace hdf5_tools {
class File {
  public:
    std::map<std::string, std::string> get_attr_map(const std::string &) const { return {}; }
    template <class T> void read(const std::string &, std::vector<T> &) const {}
};
}
