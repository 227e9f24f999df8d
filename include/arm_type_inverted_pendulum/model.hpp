// drop-in name for arm_type_inverted_pendulum/model.hpp of the reference
#pragma once
#include "cgmres_b200/models.hpp"
typedef cgmres_b200::ArmPendulumModel Model;
