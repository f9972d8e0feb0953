// TEST INFRASTRUCTURE ONLY -- the few SDL2 names app.cpp uses (see oracle/shim_app/README.md).
#pragma once
#include <cstdint>
typedef uint8_t Uint8;
typedef uint16_t Uint16;
typedef uint32_t Uint32;
typedef uint16_t SDL_AudioFormat;
struct SDL_AudioSpec {
  int freq = 0;
  SDL_AudioFormat format = 0;
  Uint8 channels = 0;
  Uint8 silence = 0;
  Uint16 samples = 0;
  Uint32 size = 0;
};
#define AUDIO_F32LSB 0x8120
#define SDL_PRESSED 1
#define SDL_BUTTON_LEFT 1
#define SDL_BUTTON_MIDDLE 2
#define SDL_BUTTON_RIGHT 3
#define SDL_BUTTON(X) (1 << ((X)-1))
#define SDL_BUTTON_LMASK SDL_BUTTON(SDL_BUTTON_LEFT)
#define SDL_BUTTON_MMASK SDL_BUTTON(SDL_BUTTON_MIDDLE)
#define SDL_BUTTON_RMASK SDL_BUTTON(SDL_BUTTON_RIGHT)
enum { KMOD_LCTRL = 0x0040, KMOD_RCTRL = 0x0080, KMOD_LALT = 0x0100, KMOD_RALT = 0x0200, KMOD_LSHIFT = 0x0001, KMOD_RSHIFT = 0x0002 };
inline int SDL_GetModState() { return 0; }
