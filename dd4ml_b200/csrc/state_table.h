// Name/offset/count table of dd4ml_state, exported so the Python side never
// hard-codes the layout.
#pragma once
#include <stddef.h>
#include <string.h>

#include "state.h"

struct dd4ml_field_desc { const char* name; int offset; int count; };

static const dd4ml_field_desc DD4ML_FIELD_TABLE[] = {
#define X(name, count) {#name, (int)(offsetof(dd4ml_state, name) / sizeof(double)), count},
    DD4ML_STATE_FIELDS(X)
#undef X
};
static const int DD4ML_NUM_FIELDS = (int)(sizeof(DD4ML_FIELD_TABLE) / sizeof(DD4ML_FIELD_TABLE[0]));
