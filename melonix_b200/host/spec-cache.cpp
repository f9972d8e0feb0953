// melonix_b200/host/spec-cache.cpp -- see spec-cache.hpp.  Replaces reference spec-cache.cpp:5-116.
#include "spec-cache.hpp"

#include <array>

SpecCache::SpecCache(Spec &aSpec, float aK, int screenWidth, double aRangeTime, std::function<int(double)> aTime2Sample)
  : spec(aSpec), k(aK), width(screenWidth), rangeTime(aRangeTime), time2Sample(std::move(aTime2Sample))
{
}

auto SpecCache::getTex(double start) -> GLuint
{
  const auto key = static_cast<int>(start * width / rangeTime); // column index (reference spec-cache.cpp:12)
  if (const auto hit = range2Tex.find(key); hit != std::end(range2Tex))
  {
    age.erase(hit->second.age);
    age.push_front(key);
    hit->second.age = std::begin(age);
    return populateTex(hit->second, key);
  }

  Column col;
  if (range2Tex.size() >= static_cast<size_t>(MaxRanges))
  {
    // recycle the least recently used column's GL texture (the reference does the same but reads
    // the list node after erasing it, spec-cache.cpp:38-40; here the key is taken first)
    const int oldestKey = age.back();
    age.pop_back();
    const auto victim = range2Tex.find(oldestKey);
    col.texture = std::move(victim->second.texture);
    range2Tex.erase(victim);
  }
  age.push_front(key);
  col.age = std::begin(age);
  const auto ins = range2Tex.emplace(key, std::move(col)).first;
  return populateTex(ins->second, key);
}

auto SpecCache::populateTex(Column &col, int key) -> GLuint
{
  const GLuint texture = col.texture.get();
  glBindTexture(GL_TEXTURE_1D, texture);
  glTexParameteri(GL_TEXTURE_1D, GL_TEXTURE_MAG_FILTER, GL_NEAREST);
  glTexParameteri(GL_TEXTURE_1D, GL_TEXTURE_MIN_FILTER, GL_NEAREST);
  if (!col.isDirty)
    return texture; // warm column: nothing to do (reference spec-cache.cpp:58-61)

  // job range = one pixel of time, warped through the markers (reference spec-cache.cpp:63-65)
  const auto start = key * rangeTime / width;
  const auto pixelSize = rangeTime / width;
  const auto rgb = spec.get().getSpecRgb(time2Sample(start), time2Sample(start + pixelSize), k);

  if (rgb.empty())
  {
    // not ready yet: 16 black texels, stay dirty (reference spec-cache.cpp:67-72)
    static const std::array<std::array<unsigned char, 3>, 16> black{};
    glTexImage1D(GL_TEXTURE_1D, 0, 3, static_cast<GLsizei>(black.size()), 0, GL_RGB, GL_UNSIGNED_BYTE, black.data());
    return texture;
  }
  col.isDirty = false;
  glTexImage1D(GL_TEXTURE_1D, 0, 3, static_cast<GLsizei>(rgb.size()), 0, GL_RGB, GL_UNSIGNED_BYTE, rgb.data());
  return texture;
}

auto SpecCache::clear() -> void
{
  range2Tex.clear();
  age.clear();
}
