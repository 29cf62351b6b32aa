#pragma once
namespace sq2_ccv_roll_pitch_msgs { struct RollPitch { double roll = 0, pitch = 0; }; }
