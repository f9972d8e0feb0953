// TEST INFRASTRUCTURE ONLY -- stand-in for mika314/sdlpp: an audio device that never calls back.
#pragma once
#include <SDL.h>
namespace sdl {
class Audio {
public:
  template <class F>
  Audio(const char *, bool, const SDL_AudioSpec *want, SDL_AudioSpec *have, int, F &&) {
    if (want && have) *have = *want;
  }
  void pause(bool) {}
  void lock() {}
  void unlock() {}
};
} // namespace sdl
