#pragma once
namespace ccv_dynamixel_msgs { struct CmdPoseByRadian { double steer_l = 0, steer_r = 0, fore = 0, rear = 0, roll = 0; }; }
