#pragma once
#include "register_macro.hpp"
