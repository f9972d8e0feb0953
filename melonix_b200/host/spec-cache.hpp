// melonix_b200/host/spec-cache.hpp -- drop-in replacement for the reference's `SpecCache`
// (spec-cache.hpp:13-39): same constructor, getTex(double) and clear().
//
// Per-pixel-column LRU of GL_TEXTURE_1D objects (<= MaxRanges).  The column's spectrum comes from
// Spec; the float -> RGB8 colour ramp (reference spec-cache.cpp:77-96) is no longer a host loop but
// the fused epilogue of the GPU kernel (Spec::getSpecRgb -> mlx_spec_batch_rgb).
#pragma once
#include "spec.hpp"
#include "texture.hpp"
#include <functional>
#include <list>
#include <unordered_map>
#include <vector>

class SpecCache
{
public:
  SpecCache(Spec &, float k, int screenWidth, double rangeTime, std::function<int(double)> time2Sample);
  auto getTex(double time) -> GLuint;
  auto clear() -> void;

private:
  struct Column
  {
    Texture texture;
    std::list<int>::iterator age;
    bool isDirty = true;
  };

  std::reference_wrapper<Spec> spec;
  float k;
  int width;
  double rangeTime;
  std::function<int(double)> time2Sample;
  std::unordered_map<int, Column> range2Tex;
  std::list<int> age; // most recently used column key first

  auto populateTex(Column &col, int key) -> GLuint;
};
