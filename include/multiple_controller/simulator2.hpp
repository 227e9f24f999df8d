// multiple_controller/simulator2.hpp of the reference is the arm plant renamed Simulator2
#pragma once
#include "cgmres_b200/models.hpp"
typedef cgmres_b200::ArmPendulumSimulator Simulator2;
