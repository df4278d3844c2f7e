// mgc_util.hpp — error plumbing shared by the C-ABI translation units
#pragma once
#include <functional>
#include <stdexcept>
#include <string>

namespace mgc {

struct ScratchExhausted : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct AlphabetError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

std::string &last_error();                       // thread-local message behind mgc_last_error()
int set_error(int code, const std::string &msg); // stores msg, returns code

}  // namespace mgc
