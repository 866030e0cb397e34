// TEST INFRASTRUCTURE: empty stand-in for <boost/enable_shared_from_this.hpp>; the reference only uses the std:: smart pointers when ROS >= 1.14.1
#pragma once
#include <memory>
