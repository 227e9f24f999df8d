// multiple_controller/model2.hpp of the reference is arm_type_inverted_pendulum/model.hpp renamed Model2
#pragma once
#include "cgmres_b200/models.hpp"
typedef cgmres_b200::ArmPendulumModel Model2;
