// TEST INFRASTRUCTURE -- stand-in for the two boost.serialization names DBoW2's BowVector.h / FeatureVector.h mention
// inside member templates that are never instantiated here (boost is absent in this image).  See cvmini.hpp.
#pragma once
namespace boost { namespace serialization {
class access {};
template <class Base, class Derived> Base& base_object(Derived& d) { return static_cast<Base&>(d); }
}}  // namespace boost::serialization
