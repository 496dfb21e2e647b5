// TEST INFRASTRUCTURE ONLY -- see pinocchio/stub.hpp
#pragma once
#include "pinocchio/stub.hpp"
