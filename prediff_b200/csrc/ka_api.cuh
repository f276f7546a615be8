// Definition of the opaque pd_ka handle (shared by api_ka.cu and api_models.cu's aligned sampling loop).
#pragma once
#include "ka.cuh"

struct pd_ka {
    pd::KANet impl;
    explicit pd_ka(const pd_ka_config& c) : impl(c) {}
};
