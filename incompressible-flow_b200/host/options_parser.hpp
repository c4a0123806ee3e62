// options_parser.hpp -- command-line parser with the call surface of mfem::OptionsParser that
// Param::add_options uses [REF src/param.cpp:15-21,115-170] [REF src/run_serial.cpp:25-39]:
// AddOption for int / double / const char* / bool (enable/disable pair), Parse, Good, PrintUsage, PrintOptions.
#pragma once
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

namespace mfem {

class OptionsParser {
 public:
  OptionsParser(int argc, char** argv) : argc_(argc), argv_(argv), error_(0), help_(false) {}
  void AddOption(int* v, const char* s, const char* l, const char* d) { opts_.push_back({INT, v, s, l, nullptr, nullptr, d}); }
  void AddOption(double* v, const char* s, const char* l, const char* d) { opts_.push_back({DOUBLE, v, s, l, nullptr, nullptr, d}); }
  void AddOption(const char** v, const char* s, const char* l, const char* d) { opts_.push_back({STRING, v, s, l, nullptr, nullptr, d}); }
  void AddOption(bool* v, const char* en_s, const char* en_l, const char* dis_s, const char* dis_l, const char* d) { opts_.push_back({BOOL, v, en_s, en_l, dis_s, dis_l, d}); }

  void Parse() {
    for (int i = 1; i < argc_; ++i) {
      const char* a = argv_[i];
      if (!std::strcmp(a, "-h") || !std::strcmp(a, "--help")) { help_ = true; return; }
      bool found = false;
      for (Opt& o : opts_) {
        const bool en = !std::strcmp(a, o.s) || !std::strcmp(a, o.l);
        const bool dis = o.type == BOOL && (!std::strcmp(a, o.ds) || !std::strcmp(a, o.dl));
        if (!en && !dis) continue;
        found = true;
        if (o.type == BOOL) { *static_cast<bool*>(o.p) = en; break; }
        if (i + 1 >= argc_) { error_ = 2; bad_ = a; return; }  // missing argument
        const char* val = argv_[++i];
        char* end = nullptr;
        if (o.type == INT) { long v = std::strtol(val, &end, 10); if (*end) { error_ = 3; bad_ = a; return; } *static_cast<int*>(o.p) = (int)v; }
        else if (o.type == DOUBLE) { double v = std::strtod(val, &end); if (*end) { error_ = 3; bad_ = a; return; } *static_cast<double*>(o.p) = v; }
        else *static_cast<const char**>(o.p) = val;
        break;
      }
      if (!found) { error_ = 1; bad_ = a; return; }  // unrecognized option
    }
  }
  bool Good() const { return !help_ && error_ == 0; }
  bool Help() const { return help_; }

  void PrintUsage(std::ostream& os) const {
    if (error_ == 1) os << "Unrecognized option: " << bad_ << "\n\n";
    else if (error_ == 2) os << "Missing argument for the last option: " << bad_ << "\n\n";
    else if (error_ == 3) os << "Wrong option format: " << bad_ << "\n\n";
    os << "Usage: " << argv_[0] << " [options] ...\nOptions:\n   -h, --help\n\tPrint this help message and exit.\n";
    for (const Opt& o : opts_) {
      os << "   " << o.s << (o.type == BOOL ? "" : type_name(o.type)) << ", " << o.l << (o.type == BOOL ? "" : type_name(o.type));
      if (o.type == BOOL) os << ", " << o.ds << ", " << o.dl;
      os << ", current value: ";
      print_value(os, o, true);
      os << "\n\t" << o.d << "\n";
    }
  }
  void PrintOptions(std::ostream& os) const {
    os << "Options used:\n";
    for (const Opt& o : opts_) {
      os << "   ";
      if (o.type == BOOL) os << (*static_cast<bool*>(o.p) ? o.l : o.dl);
      else { os << o.l << " "; print_value(os, o, false); }
      os << "\n";
    }
  }

 private:
  enum Type { INT, DOUBLE, STRING, BOOL };
  struct Opt { Type type; void* p; const char *s, *l, *ds, *dl, *d; };
  static const char* type_name(Type t) { return t == INT ? " <int>" : (t == DOUBLE ? " <double>" : " <string>"); }
  static void print_value(std::ostream& os, const Opt& o, bool usage) {
    switch (o.type) {
      case INT: os << *static_cast<int*>(o.p); break;
      case DOUBLE: os << *static_cast<double*>(o.p); break;
      case STRING: os << *static_cast<const char**>(o.p); break;
      case BOOL: os << (*static_cast<bool*>(o.p) ? o.s : o.ds); (void)usage; break;
    }
  }
  int argc_; char** argv_; int error_; bool help_; std::string bad_;
  std::vector<Opt> opts_;
};

}  // namespace mfem
