#pragma once
#include "srst_device.cuh"
