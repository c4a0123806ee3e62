#include "utilities.hpp"

#include <cstdio>
#include <fstream>
#include <vector>

#include "config.hpp"

// Same contract as the reference: the element size is inferred from file length / n_values; float32
// values are widened to double one by one [REF src/utilities.cpp:43-51].  Unlike the reference, float64
// files are accepted too (the reference's own gcount() check rejects them) and files larger
// than 2 GiB do not overflow.
void read_binary(const char* filename, int n_values, double* values) {
  std::ifstream in(filename, std::ios::binary | std::ios::ate);
  MFEM_VERIFY(in, "File '" + std::string(filename) + "' can't be opened");
  const long long length = (long long)in.tellg();
  MFEM_VERIFY(n_values > 0 && length % n_values == 0,
              "The number of bytes in the file '" + std::string(filename) +
                  "' is not divisible by the number of elements " + d2s(n_values));
  const long long size_value = length / n_values;
  in.seekg(0, std::ios::beg);
  if (size_value == (long long)sizeof(double)) {
    in.read(reinterpret_cast<char*>(values), (std::streamsize)n_values * sizeof(double));
    MFEM_VERIFY(in.gcount() == (std::streamsize)n_values * (std::streamsize)sizeof(double),
                "The number of successfully read elements is different from the expected one");
  } else if (size_value == (long long)sizeof(float)) {
    std::vector<float> buf(n_values);
    in.read(reinterpret_cast<char*>(buf.data()), (std::streamsize)n_values * sizeof(float));
    MFEM_VERIFY(in.gcount() == (std::streamsize)n_values * (std::streamsize)sizeof(float),
                "The number of successfully read elements is different from the expected one");
    for (int i = 0; i < n_values; ++i) values[i] = buf[i];
  } else {
    MFEM_ABORT("Unknown size of an element (" + d2s(size_value) + ") in bytes. Expected one is either sizeof(float) = " +
               d2s(sizeof(float)) + ", or sizeof(double) = " + d2s(sizeof(double)));
  }
}

void write_binary(const char* filename, int n_values, const double* values) {
  std::ofstream out(filename, std::ios::binary);
  MFEM_VERIFY(out, "File '" + std::string(filename) + "' can't be opened");
  std::vector<float> buf(n_values);
  for (int i = 0; i < n_values; ++i) buf[i] = (float)values[i];
  out.write(reinterpret_cast<const char*>(buf.data()), (std::streamsize)n_values * sizeof(float));
}
