// drop-in name for arm_type_inverted_pendulum/simulator.hpp of the reference
#pragma once
#include "cgmres_b200/models.hpp"
typedef cgmres_b200::ArmPendulumSimulator Simulator;
