// TEST INFRASTRUCTURE ONLY -- stand-in for mika314/ser (project files are out of scope).
#pragma once
class IStrm {
public:
  IStrm(const char *, const char *) {}
};
