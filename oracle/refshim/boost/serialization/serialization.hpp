// oracle/refshim -- TEST INFRASTRUCTURE ONLY: stands in for the Boost.Serialization header of the same name.
// DBoW2's BowVector.h / FeatureVector.h only mention it inside never-instantiated `serialize` templates.
#pragma once
namespace boost {
namespace serialization {
class access;
template <class Base, class Derived>
Base& base_object(Derived& d) { return static_cast<Base&>(d); }
}  // namespace serialization
}  // namespace boost
