// TEST INFRASTRUCTURE ONLY -- stand-in for mika314/ser: the property list expands to nothing.
#pragma once
#define SER_DEF_PROPS()
