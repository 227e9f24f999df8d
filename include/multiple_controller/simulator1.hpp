// multiple_controller/simulator1.hpp of the reference: the msd plant with t_end = 10 (simulator1.hpp:6)
#pragma once
#include "cgmres_b200/models.hpp"
struct Simulator1 : cgmres_b200::MassSpringDamperSimulator {
  static constexpr double t_end = 10;
};
