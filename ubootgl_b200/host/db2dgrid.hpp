// db2dgrid.hpp -- host-side grid containers of the B200 drop-in.
//
// Same public surface as the reference's Single2DGrid / DoubleBuffered2DGrid
// (te42kyfo/ubootgl db2dgrid.hpp:12-50, 52-109): unpadded row-major fp32,
// idx = y*width + x, public width/height, operator()(x,y), data(), fill(),
// f()/b()/swap()/back_data().  Here they are the HOST MIRRORS of fields that
// live on the GPU: every mutable access raises a dirty flag so that
// Simulation::step() uploads exactly the grids the game touched (SURVEY.md 8b:
// the reference's real interface is its public data members).
#pragma once
#include "ubgl_vec.hpp"
#include <algorithm>
#include <cassert>
#include <cstddef>
#include <functional>
#include <utility>
#include <vector>

namespace ubgl_host {
// row-major fp32 storage with a "host wrote to me" flag (dirty: the next step() uploads it) and
// a "device is newer" flag (stale: set by Simulation in SyncMode::RESIDENT, where step() does not
// download; the first host access then pulls the field from the device, so a mirror the game
// reads or edits is never older than the device state it is about to overwrite).
class MirrorStore {
public:
  using Pull = std::function<void(float *)>;
  MirrorStore() = default;
  MirrorStore(int w, int h) : cells_((size_t)w * h, 0.0f) {}
  // value semantics like std::vector<float> (the reference copy-assigns grids, simulation.hpp:38-43):
  // a copy is a plain host grid (no device behind it); assignment keeps the destination's device link
  MirrorStore(const MirrorStore &o) : cells_((o.fresh(), o.cells_)), dirty_(true) {}
  MirrorStore(MirrorStore &&o) : cells_((o.fresh(), std::move(o.cells_))), dirty_(true) {}
  MirrorStore &operator=(const MirrorStore &o) {
    if (this != &o) {
      o.fresh();
      cells_ = o.cells_;
      dirty_ = true;
      stale_ = false;
    }
    return *this;
  }
  MirrorStore &operator=(MirrorStore &&o) {
    if (this != &o) {
      o.fresh();
      cells_ = std::move(o.cells_);
      dirty_ = true;
      stale_ = false;
    }
    return *this;
  }
  float *rw() {
    fresh();
    dirty_ = true;
    return cells_.data();
  }
  const float *ro() const {
    fresh();
    return cells_.data();
  }
  float *raw() { return cells_.data(); } // library-side access: neither marks nor pulls
  size_t size() const { return cells_.size(); }
  bool dirty() const { return dirty_; }
  void clean() { dirty_ = false; }
  void touch() { dirty_ = true; }
  // ---- used by Simulation only ----
  void set_pull(Pull p) { pull_ = std::move(p); }
  void mark_stale() { stale_ = pull_ != nullptr; }
  void mark_fresh() { stale_ = false; }
  bool stale() const { return stale_; }

private:
  void fresh() const {
    if (!stale_) return;
    stale_ = false;
    pull_(const_cast<float *>(cells_.data()));
  }
  std::vector<float> cells_;
  bool dirty_ = true; // a fresh grid has never been uploaded
  mutable bool stale_ = false;
  Pull pull_;
};
} // namespace ubgl_host

class Single2DGrid {
public:
  Single2DGrid() = default;
  Single2DGrid(int w, int h) : width(w), height(h), s_(w, h) {}

  int idx(int x, int y) const { return y * width + x; }
  float *data() { return s_.rw(); }
  const float *data() const { return s_.ro(); }
  float &operator()(int x, int y) {
    assert(inside(x, y));
    return s_.rw()[idx(x, y)];
  }
  float operator()(int x, int y) const {
    assert(inside(x, y));
    return s_.ro()[idx(x, y)];
  }
  float &operator()(glm::ivec2 c) { return (*this)(c.x, c.y); }
  float operator()(glm::ivec2 c) const { return (*this)(c.x, c.y); }
  void fill(float v) { std::fill_n(s_.rw(), s_.size(), v); }

  int width = 0, height = 0;

  // ---- mirror protocol (used by MG / Simulation, not by the game) ----
  ubgl_host::MirrorStore &mirror() { return s_; }
  const ubgl_host::MirrorStore &mirror() const { return s_; }

private:
  bool inside(int x, int y) const { return x >= 0 && y >= 0 && x < width && y < height; }
  ubgl_host::MirrorStore s_;
};

class DoubleBuffered2DGrid {
public:
  DoubleBuffered2DGrid() = default;
  DoubleBuffered2DGrid(int w, int h) : width(w), height(h) {
    s_[0] = ubgl_host::MirrorStore(w, h);
    s_[1] = ubgl_host::MirrorStore(w, h);
  }
  DoubleBuffered2DGrid(const DoubleBuffered2DGrid &) = delete; // db2dgrid.hpp:61
  DoubleBuffered2DGrid &operator=(const DoubleBuffered2DGrid &) = default;
  DoubleBuffered2DGrid &operator=(DoubleBuffered2DGrid &&) = default;

  int idx(int x, int y) const { return y * width + x; }
  void swap() { front_ ^= 1; }

  float &f(int x, int y) { return s_[front_].rw()[idx(x, y)]; }
  float &b(int x, int y) { return s_[front_ ^ 1].rw()[idx(x, y)]; }
  float f(int x, int y) const { return s_[front_].ro()[idx(x, y)]; }
  float b(int x, int y) const { return s_[front_ ^ 1].ro()[idx(x, y)]; }
  float &f(glm::ivec2 c) { return f(c.x, c.y); }
  float &b(glm::ivec2 c) { return b(c.x, c.y); }
  float f(glm::ivec2 c) const { return f(c.x, c.y); }
  float b(glm::ivec2 c) const { return b(c.x, c.y); }
  float &operator()(int x, int y) { return f(x, y); }
  float operator()(int x, int y) const { return f(x, y); }
  float &operator()(glm::ivec2 c) { return f(c.x, c.y); }
  float operator()(glm::ivec2 c) const { return f(c.x, c.y); }

  void copyFrontToBack() {
    std::copy_n(s_[front_].ro(), s_[front_].size(), s_[front_ ^ 1].rw());
  }
  float *data() { return s_[front_].rw(); }
  const float *data() const { return s_[front_].ro(); }
  float *back_data() { return s_[front_ ^ 1].rw(); }

  int width = 0, height = 0;

  ubgl_host::MirrorStore &front_mirror() { return s_[front_]; }
  ubgl_host::MirrorStore &back_mirror() { return s_[front_ ^ 1]; }
  ubgl_host::MirrorStore &store(int k) { return s_[k]; }
  int front_index() const { return front_; }

private:
  int front_ = 0;
  ubgl_host::MirrorStore s_[2];
};
