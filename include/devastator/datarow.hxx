// devastator/datarow.hxx - a row of name=value assignments, printable as python kwargs
// (reference surface: src/devastator/datarow.hxx:22-137; used by bench/util/report.cxx).
#ifndef DEVA_B200_DATAROW_HXX
#define DEVA_B200_DATAROW_HXX
#include <functional>
#include <iosfwd>
#include <map>
#include <sstream>
#include <string>
#include <utility>

namespace deva {
  class datarow {
    struct cell {
      bool is_y = false;     // dependent (y) vs independent (x) variable
      bool is_num = true;
      double num = 0.0;
      std::string str;
      cell() = default;
      cell(double v, bool y): is_y(y), is_num(true), num(v) {}
      cell(int v, bool y): is_y(y), is_num(true), num(v) {}
      cell(bool v, bool y): is_y(y), is_num(true), num(v) {}
      cell(std::string v, bool y): is_y(y), is_num(false), str(std::move(v)) {}
      cell(const char *v, bool y): is_y(y), is_num(false), str(v ? v : "") {}
      bool operator==(cell const &o) const {
        return is_y == o.is_y && is_num == o.is_num && (is_num ? num == o.num : str == o.str);
      }
      void to_python(std::ostream &o) const;
    };
    std::map<std::string, cell> m_;
    friend struct std::hash<datarow>;
  public:
    datarow() = default;
    template<typename T>
    datarow(std::string name, T val, bool y_not_x) { m_.emplace(std::move(name), cell(std::move(val), y_not_x)); }

    bool operator==(datarow const &o) const { return m_ == o.m_; }
    bool operator!=(datarow const &o) const { return !(m_ == o.m_); }

    void to_python_kwargs(bool x_not_y, std::ostream &o, bool leading_comma);
    void xs_to_python_kwargs(std::ostream &o, bool leading_comma) { to_python_kwargs(true, o, leading_comma); }
    void ys_to_python_kwargs(std::ostream &o, bool leading_comma) { to_python_kwargs(false, o, leading_comma); }
    static datarow from_python_kwargs(std::istream &i);
    static datarow from_python_kwargs(std::string const &s) { std::istringstream in(s); return from_python_kwargs(in); }

    template<typename T> static datarow x(std::string name, T val) { return datarow(std::move(name), std::move(val), false); }
    template<typename T> static datarow y(std::string name, T val) { return datarow(std::move(name), std::move(val), true); }

    friend datarow& operator&=(datarow &a, datarow b);
    friend datarow operator&(datarow a, datarow b);
    datarow prefix_all(std::string prefix);
  };
  datarow& operator&=(datarow &a, datarow b);
  datarow operator&(datarow a, datarow b);
}
namespace std {
  template<> struct hash<deva::datarow> { std::size_t operator()(deva::datarow const &row) const; };
}
#endif
