// Stub standing in for the reference's vendored fast5.hpp so that the reference's
// mapping path (src/mapper.cpp etc.) can be compiled WITHOUT HDF5.  Test infrastructure only.
#pragma once
#include <fast5/hdf5_tools.hpp>
