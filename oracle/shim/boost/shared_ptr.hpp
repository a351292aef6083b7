/* TEST INFRASTRUCTURE ONLY -- boost::shared_ptr stand-in (alias of std::shared_ptr). */
#ifndef ORACLE_SHIM_BOOST_SHARED_PTR_HPP
#define ORACLE_SHIM_BOOST_SHARED_PTR_HPP
#include <memory>
namespace boost {
template <class T> using shared_ptr = std::shared_ptr<T>;
}
#endif
