#pragma once
namespace gazebo_msgs { struct GetLinkState {}; }
