// boost/filesystem.hpp stand-in -- TEST INFRASTRUCTURE ONLY (see oracle/Makefile).
//
// Boost is not installed in this image.  The reference driver uses three things from it (src/main.cpp:62-68):
// construction of a path from a C string or std::string, operator/ and c_str().  This header provides those
// with Boost.Filesystem's joining rule (a separator is inserted unless the left side is empty or already
// ends with one) so the UNMODIFIED src/main.cpp compiles where it lies.  Nothing in the product includes it.
#ifndef PFS_COMPAT_BOOST_FILESYSTEM_HPP
#define PFS_COMPAT_BOOST_FILESYSTEM_HPP

#include <string>

namespace boost {
namespace filesystem {

class path {
public:
    path() {}
    path(const char *s) : s_(s ? s : "") {}
    path(const std::string &s) : s_(s) {}
    const char *c_str() const { return s_.c_str(); }
    const std::string &string() const { return s_; }
    path &operator/=(const path &rhs)
    {
        if (!s_.empty() && s_.back() != '/' && !rhs.s_.empty() && rhs.s_.front() != '/') s_ += '/';
        s_ += rhs.s_;
        return *this;
    }

private:
    std::string s_;
};

inline path operator/(const path &lhs, const path &rhs)
{
    path out(lhs);
    out /= rhs;
    return out;
}

}  // namespace filesystem
}  // namespace boost
#endif
