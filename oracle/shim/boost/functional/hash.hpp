// TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Empty on purpose: boost::hash_value for the bitset
// stand-in lives in ../dynamic_bitset.hpp (the reference includes this header at tesseract.cc:18).
#ifndef ORACLE_SHIM_BOOST_FUNCTIONAL_HASH_HPP
#define ORACLE_SHIM_BOOST_FUNCTIONAL_HASH_HPP
#endif
