// TEST INFRASTRUCTURE: stand-in for <ros/common.h>
#pragma once
#define ROS_VERSION_MINIMUM(a, b, c) 1
