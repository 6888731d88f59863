// xstrom.hpp -- the exception type every host-side failure surfaces as
// (mirrors reference src/xstrom.hpp:3-14; engine error codes are converted by Likelihood).
#pragma once
#include <exception>
#include <string>

namespace strom {

class XStrom : public std::exception {
  public:
    XStrom() noexcept {}
    explicit XStrom(const std::string& s) noexcept : _msg(s) {}
    const char* what() const noexcept override { return _msg.c_str(); }

  private:
    std::string _msg;
};

}  // namespace strom
