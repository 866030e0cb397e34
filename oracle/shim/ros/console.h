// TEST INFRASTRUCTURE: stand-in for <ros/console.h>: logging is dropped (arguments are not evaluated)
#pragma once
#include <cstdio>

#include "common.h"
#include <iostream>
#include <sstream>
#define ROS_DEBUG(...) do { } while (0)
#define ROS_INFO(...) do { } while (0)
#define ROS_WARN(...) do { } while (0)
#define ROS_ERROR(...) do { } while (0)
#define ROS_FATAL(...) do { } while (0)
#define ROS_DEBUG_STREAM(x) do { } while (0)
#define ROS_INFO_STREAM(x) do { } while (0)
#define ROS_WARN_STREAM(x) do { } while (0)
#define ROS_ERROR_STREAM(x) do { } while (0)
#define ROS_WARN_THROTTLE(...) do { } while (0)
#define ROS_DEBUG_THROTTLE(...) do { } while (0)
