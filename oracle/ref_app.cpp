/* oracle/ref_app.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Drives the reference's own application class `App` (compiled UNMODIFIED from
 * /root/reference/app.cpp, spec-cache.cpp, save-wav.cpp and spec.cpp by oracle/Makefile target `ref`,
 * against the no-op UI / audio / codec headers of oracle/shim_app) so that the oracle's restatements of
 *   App::preproc grain segmentation      app.cpp:156-235      (oracle/grain_ref.c)
 *   App::process / App::exportWav        app.cpp:294-345, 1194-1215 + saveWav, save-wav.cpp:17-48
 *   sample2Time / time2Sample / duration / time2PitchBend     app.cpp:1020-1122
 *   App::calcPicks / getMinMaxFromRange  app.cpp:347-426      (oracle/picks_ref.c)
 *   SpecCache::populateTex colour ramp   spec-cache.cpp:77-96 (oracle/colormap_ref.c)
 * are pinned against the reference code itself.  The members involved are private; this translation
 * unit (and only this one) reads the class definition with `private` spelled `public`, which changes
 * neither layout nor symbol names.  FileOpen / FileSaveAs (ImGui dialogs, out of scope) get empty
 * member definitions here instead of compiling file-open.cpp / file-save-as.cpp.
 */
#include <algorithm>
#include <array>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <functional>
#include <list>
#include <map>
#include <memory>
#include <span>
#include <string>
#include <thread>
#include <tuple>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include <SDL.h>
#include <SDL_opengl.h>
#include <fftw3.h>
#include <imgui/imgui.h>
#include <sdlpp/sdlpp.hpp>
#include <ser/macro.hpp>

#define private public
#include <app.hpp> /* the reference header, found via -I /root/reference */
#undef private

auto FileOpen::draw() -> bool { return false; }
auto FileOpen::getSelectedFile() const -> std::filesystem::path { return {}; }
FileSaveAs::FileSaveAs(std::string name) : dialogName(std::move(name)) { fileName.fill(0); }
auto FileSaveAs::draw() -> bool { return false; }
auto FileSaveAs::getSelectedFile() const -> std::string { return {}; }

struct mlxo_ref_marker {
  int sample;
  double note, dTime, pitchBend;
};

extern "C" {

/* A fresh App holding `wav` at `sampleRate` with `markers` (sorted by sample), after App::preproc():
 * grains, picks and the Spec are built by the reference's own code. */
static std::string g_last_error;
const char *mlxo_ref_app_last_error() { return g_last_error.c_str(); }

void *mlxo_ref_app_create(const float *wav, long long n, int sampleRate, const mlxo_ref_marker *m, int nm) {
  auto *a = new App();
  a->wavData.assign(wav, wav + n);
  a->sampleRate = sampleRate;
  for (int i = 0; i < nm; ++i) {
    Marker mk;
    mk.sample = m[i].sample;
    mk.note = m[i].note;
    mk.dTime = m[i].dTime;
    mk.pitchBend = m[i].pitchBend;
    a->markers.push_back(mk);
  }
  try {
    a->preproc();
  } catch (const std::exception &e) { /* only the drop-in build can throw: its Spec needs a B200 */
    g_last_error = e.what();
    delete a;
    return nullptr;
  }
  return a;
}

void mlxo_ref_app_destroy(void *h) { delete static_cast<App *>(h); }

/* grains of App::preproc: (start, length) = key and span size of App::grains */
int mlxo_ref_app_grains(void *h, int *g_start, int *g_len, int cap) {
  auto *a = static_cast<App *>(h);
  int i = 0;
  for (const auto &g : a->grains) {
    if (i < cap) {
      g_start[i] = g.first;
      g_len[i] = static_cast<int>(std::get<0>(g.second).size());
    }
    ++i;
  }
  return i;
}

/* App::picks, level after level; level_off[levels + 1] in pairs; returns the number of levels */
int mlxo_ref_app_picks(void *h, float *pairs, long long cap_pairs, long long *level_off) {
  auto *a = static_cast<App *>(h);
  long long off = 0;
  int l = 0;
  for (const auto &lvl : a->picks) {
    level_off[l++] = off;
    for (const auto &p : lvl) {
      if (off < cap_pairs) {
        pairs[2 * off] = p.first;
        pairs[2 * off + 1] = p.second;
      }
      ++off;
    }
  }
  level_off[l] = off;
  return l;
}

void mlxo_ref_app_minmax(void *h, const int *start_end, int count, float *out) {
  auto *a = static_cast<App *>(h);
  for (int i = 0; i < count; ++i) {
    const auto mm = a->getMinMaxFromRange(start_end[2 * i], start_end[2 * i + 1]);
    out[2 * i] = mm.first;
    out[2 * i + 1] = mm.second;
  }
}

/* the warp maps, memo caches cleared before every call (App::invalidateCache, app.cpp:840-852) */
double mlxo_ref_app_sample2time(void *h, int s) {
  auto *a = static_cast<App *>(h);
  a->invalidateCache();
  return a->sample2Time(s);
}
int mlxo_ref_app_time2sample(void *h, double t) {
  auto *a = static_cast<App *>(h);
  a->invalidateCache();
  return a->time2Sample(t);
}
double mlxo_ref_app_duration(void *h) {
  auto *a = static_cast<App *>(h);
  a->invalidateCache();
  return a->duration();
}
float mlxo_ref_app_time2pitchbend(void *h, double t) {
  auto *a = static_cast<App *>(h);
  a->invalidateCache();
  return a->time2PitchBend(t);
}

/* App::exportWav into `path` (the reference's own loop, conversion and saveWav). */
void mlxo_ref_app_export(void *h, const char *path) {
  auto *a = static_cast<App *>(h);
  a->invalidateCache();
  a->exportWav(path);
}

/* The float samples exportWav converts: its cursor loop (app.cpp:1201-1207) around the reference's
 * App::process.  Returns the length; writes up to cap samples. */
long long mlxo_ref_app_render(void *h, float *pcm, long long cap) {
  auto *a = static_cast<App *>(h);
  a->invalidateCache();
  std::vector<float> out;
  for (auto cursor = 0.;;) {
    const auto dt = a->process(cursor, out);
    if (dt <= 0.) break;
    cursor += dt;
  }
  const long long n = static_cast<long long>(out.size());
  std::memcpy(pcm, out.data(), sizeof(float) * static_cast<size_t>(std::min(n, cap)));
  return n;
}

/* One spectrogram column through the reference's SpecCache (spec-cache.cpp:10-110): polls getTex until
 * the Spec worker has delivered, then returns the texels glTexImage1D received ([width][3]).
 * Returns the texel count (SpectrSize / 2) or < 0 on time-out. */
int mlxo_ref_app_speccache_column(void *h, float k, int screenWidth, double rangeTime, double time,
                                  unsigned char *rgb, int cap_texels) {
  auto *a = static_cast<App *>(h);
  SpecCache cache(*a->spec, k, screenWidth, rangeTime, [a](double t) { return a->time2Sample(t); });
  for (int tries = 0; tries < 20000; ++tries) {
    mlx_gl_capture().width = 0;
    cache.getTex(time);
    const MlxGlCapture &c = mlx_gl_capture();
    if (c.width > 16) {
      std::memcpy(rgb, c.rgb.data(), static_cast<size_t>(std::min(c.width, cap_texels)) * 3);
      return c.width;
    }
    std::this_thread::sleep_for(std::chrono::milliseconds(1));
  }
  return -1;
}
}
