// src/ds.h -- drop-in for the reference's src/ds.h, for the unchanged qtip_parse.cpp.
//
// Provides the two containers the caller uses (reference src/ds.h:9-395 EList, :400-491
// ReservoirSampledEList) with the same method names and semantics, on top of std::vector.
// Elements are real values here (templates own std::strings), so growth needs none of the
// reference's memset-the-old-buffer trick (ds.h:366).
#ifndef QTIP_B200_DS_H
#define QTIP_B200_DS_H

#include <cassert>
#include <cstddef>
#include <algorithm>
#include <vector>
#include "rnglib.hpp"

template <typename T, int S = 128>
class EList {
public:
	EList() { v_.reserve(S); }
	size_t size() const { return v_.size(); }
	size_t capacity() const { return v_.capacity(); }
	bool empty() const { return v_.empty(); }
	void ensure(size_t n) { v_.reserve(v_.size() + n); }
	void reserveExact(size_t n) { v_.reserve(n); }
	void push_back(const T &el) { v_.push_back(el); }
	void expand() { v_.emplace_back(); }          // new, default-constructed element (ds.h:72-76)
	void pop_back() { v_.pop_back(); }
	void clear() { v_.clear(); }
	void resize(size_t n) { v_.resize(n); }
	void resizeExact(size_t n) { v_.resize(n); }
	void fill(const T &x) { std::fill(v_.begin(), v_.end(), x); }
	void erase(size_t i) { v_.erase(v_.begin() + i); }
	void insert(const T &el, size_t i) { v_.insert(v_.begin() + i, el); }
	T &back() { assert(!v_.empty()); return v_.back(); }
	const T &back() const { assert(!v_.empty()); return v_.back(); }
	T &front() { assert(!v_.empty()); return v_.front(); }
	const T &front() const { assert(!v_.empty()); return v_.front(); }
	T &operator[](size_t i) { assert(i < v_.size()); return v_[i]; }
	const T &operator[](size_t i) const { assert(i < v_.size()); return v_[i]; }
	T &get(size_t i) { return (*this)[i]; }
	const T &get(size_t i) const { return (*this)[i]; }
	T *ptr() { return v_.data(); }
	const T *ptr() const { return v_.data(); }

private:
	std::vector<T> v_;
};

// Reservoir sampler (reference ds.h:400-491).  add_part1 consumes one r4_uni_01 draw per item
// once the reservoir is full (ds.h:437-446) -- a side effect on the global generator that
// replay parity depends on -- and returns the slot index the caller tests
// against k().
template <typename T>
class ReservoirSampledEList {
public:
	explicit ReservoirSampledEList(size_t k) : k_(k), n_(0) {}

	void add(const T &t) {
		n_++;
		if (list_.size() < k_) {
			list_.push_back(t);
		} else {
			size_t j = std::min((size_t)(r4_uni_01() * n_), n_ - 1);
			if (j < list_.size()) list_[j] = t;
		}
	}

	size_t add_part1() {
		n_++;
		if (list_.size() < k_) {
			list_.expand();
			return list_.size() - 1;
		}
		return std::min((size_t)(r4_uni_01() * n_), n_ - 1);
	}

	size_t size() const { return n_; }   // items offered, not items kept
	bool empty() const { return n_ == 0; }
	size_t k() const { return k_; }
	const EList<T> &list() const { return list_; }
	EList<T> &list() { return list_; }

private:
	size_t k_, n_;
	EList<T> list_;
};

#endif
