#pragma once
#include <ros/ros.h>
namespace std_msgs { struct Float64 { double data = 0.0; }; }
