// TEST INFRASTRUCTURE: empty stand-in for <boost/concept_check.hpp>; the reference only uses the std:: smart pointers when ROS >= 1.14.1
#pragma once
#include <memory>
