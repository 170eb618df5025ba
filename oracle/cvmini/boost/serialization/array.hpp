// TEST INFRASTRUCTURE -- stand-in for boost/serialization/array.hpp: MapPoint.h mentions make_array inside a member
// template that is never instantiated here (boost is absent in this image).
#pragma once
#include <cstddef>
namespace boost { namespace serialization {
template <class T> struct array_wrapper { T* p; std::size_t n; };
template <class T> array_wrapper<T> make_array(T* p, std::size_t n) { return array_wrapper<T>{p, n}; }
}}  // namespace boost::serialization
