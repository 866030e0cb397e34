// TEST INFRASTRUCTURE: empty stand-in for <boost/shared_ptr.hpp>; the reference only uses the std:: smart pointers when ROS >= 1.14.1
#pragma once
#include <memory>
