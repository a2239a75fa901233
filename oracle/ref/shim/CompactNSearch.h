// API-compatible stand-in for the neighbour-search library the reference links against.
//
// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// The reference uses a *fork* of InteractiveComputerGraphics/CompactNSearch as an
// un-vendored git submodule (vendor/CompactNSearch is empty in the snapshot; pinned commit
// unknown; premake5.lua:64,80).  This header restates the published behaviour of that
// library for exactly the call sites the reference has:
//   Dataset.h:55,58        members, default-constructible + copyable
//   Dataset.cpp:16-17      NeighborhoodSearch(radius)
//   Dataset.cpp:54-62      add_point_set / z_sort / point_set(i).sort_field / update_point_sets
//   Dataset.cpp:128        find_neighbors_box   (fork-only, semantics chosen below)
//   Dataset.cpp:277,287    find_neighbors(point, out)
//
// Published algorithm restated (upstream CompactNSearch, Real = float):
//   * hash cell size = r; per-axis cell index  x >= 0 ? (int)(x/r) : (int)(x/r) - 1,
//     evaluated as (int)(inv_r * x) with inv_r = 1/r in Real;
//   * a point query visits the 27 cells around the query cell (dj, dk, dl nested, each
//     -1..1) and keeps id iff (dx*dx + dy*dy) + dz*dz < r*r, accumulated left to right;
//   * result order = cell traversal order, then insertion order inside a cell; after
//     update_point_sets() insertion order is ascending point index;
//   * z_sort() orders points by the Morton code of their cell (21 bits per axis of
//     key - INT_MIN, x in the lowest bit); sort_field() applies that permutation to a
//     user array.  Ties keep their original order here (std::stable_sort) so that the
//     permutation is reproducible -- upstream uses an unstable sort ("parity unpinned").
//   * find_neighbors_box is not upstream.  The only evidence is its use (query = centre of
//     a DensityGrid cell, radius = cell width, result = "number of particles in cells").
//     g_box_mode selects the reading:
//       0 (default)  axis-aligned box of half-width r/2 around the query, half-open
//                    [c - r/2, c + r/2)  == "the particles inside that cell";
//       1            every point stored in the 27 hash cells, no distance test.
#pragma once

#include <algorithm>
#include <climits>
#include <cstddef>
#include <cstdint>
#include <numeric>
#include <unordered_map>
#include <vector>

namespace CompactNSearch
{
using Real = float;

inline int& box_mode()
{
	static int g_box_mode = 0;
	return g_box_mode;
}

struct CellKey
{
	int k[3];
	bool operator==(CellKey const& o) const { return k[0] == o.k[0] && k[1] == o.k[1] && k[2] == o.k[2]; }
};

struct CellKeyHasher
{
	std::size_t operator()(CellKey const& c) const
	{
		return std::size_t(73856093u * uint32_t(c.k[0])) ^ std::size_t(19349663u * uint32_t(c.k[1])) ^
			std::size_t(83492791u * uint32_t(c.k[2]));
	}
};

class NeighborhoodSearch;

class PointSet
{
public:
	std::size_t n_points() const { return m_n; }

	// permutes `lst` (one element per point) into z-sort order
	template <typename T>
	void sort_field(T* lst) const
	{
		if (m_sort_table.empty()) return;
		std::vector<T> tmp(lst, lst + m_sort_table.size());
		for (std::size_t i = 0; i < m_sort_table.size(); i++)
			lst[i] = tmp[m_sort_table[i]];
	}

	std::vector<unsigned> const& sort_table() const { return m_sort_table; }

private:
	friend class NeighborhoodSearch;
	Real const* m_x = nullptr;
	std::size_t m_n = 0;
	std::vector<unsigned> m_sort_table;
};

class NeighborhoodSearch
{
public:
	NeighborhoodSearch(Real r = Real(1), bool /*erase_empty_cells*/ = false) :
		m_r(r), m_r2(r * r), m_inv(Real(1) / r) {}

	unsigned add_point_set(Real const* x, std::size_t n, bool = true, bool = true, bool = true)
	{
		PointSet ps;
		ps.m_x = x;
		ps.m_n = n;
		m_sets.push_back(ps);
		m_ready = false;
		return unsigned(m_sets.size() - 1);
	}

	PointSet& point_set(unsigned i) { return m_sets[i]; }
	PointSet const& point_set(unsigned i) const { return m_sets[i]; }

	void z_sort()
	{
		for (PointSet& ps : m_sets)
		{
			std::vector<uint64_t> code(ps.m_n);
			for (std::size_t i = 0; i < ps.m_n; i++)
				code[i] = morton(cell_of(ps.m_x + 3 * i));
			ps.m_sort_table.resize(ps.m_n);
			std::iota(ps.m_sort_table.begin(), ps.m_sort_table.end(), 0u);
			std::stable_sort(ps.m_sort_table.begin(), ps.m_sort_table.end(),
							 [&](unsigned a, unsigned b) { return code[a] < code[b]; });
		}
		m_ready = false;
	}

	void update_point_sets() { rebuild(); }

	void find_neighbors(Real const* x, std::vector<std::vector<unsigned>>& out)
	{
		if (!m_ready) rebuild();
		out.clear();
		out.resize(m_sets.size());
		CellKey const c = cell_of(x);
		for (int dj = -1; dj <= 1; dj++)
			for (int dk = -1; dk <= 1; dk++)
				for (int dl = -1; dl <= 1; dl++)
				{
					auto it = m_map.find(CellKey{ { c.k[0] + dj, c.k[1] + dk, c.k[2] + dl } });
					if (it == m_map.end()) continue;
					for (Entry const& e : it->second)
					{
						Real const* xb = m_sets[e.set].m_x + 3 * std::size_t(e.id);
						Real const d0 = x[0] - xb[0], d1 = x[1] - xb[1], d2 = x[2] - xb[2];
						Real const l2 = d0 * d0 + d1 * d1 + d2 * d2;
						if (l2 < m_r2) out[e.set].push_back(e.id);
					}
				}
	}

	void find_neighbors_box(Real const* x, std::vector<std::vector<unsigned>>& out)
	{
		if (!m_ready) rebuild();
		out.clear();
		out.resize(m_sets.size());
		CellKey const c = cell_of(x);
		Real const half = Real(0.5) * m_r;
		Real const lo[3] = { x[0] - half, x[1] - half, x[2] - half };
		Real const hi[3] = { x[0] + half, x[1] + half, x[2] + half };
		for (int dj = -1; dj <= 1; dj++)
			for (int dk = -1; dk <= 1; dk++)
				for (int dl = -1; dl <= 1; dl++)
				{
					auto it = m_map.find(CellKey{ { c.k[0] + dj, c.k[1] + dk, c.k[2] + dl } });
					if (it == m_map.end()) continue;
					for (Entry const& e : it->second)
					{
						Real const* xb = m_sets[e.set].m_x + 3 * std::size_t(e.id);
						bool const inside = box_mode() == 1 ||
							(xb[0] >= lo[0] && xb[0] < hi[0] && xb[1] >= lo[1] && xb[1] < hi[1] &&
							 xb[2] >= lo[2] && xb[2] < hi[2]);
						if (inside) out[e.set].push_back(e.id);
					}
				}
	}

	Real radius() const { return m_r; }

private:
	struct Entry { unsigned set; unsigned id; };

	CellKey cell_of(Real const* x) const
	{
		CellKey c;
		for (int i = 0; i < 3; i++)
		{
			int const t = static_cast<int>(m_inv * x[i]);
			c.k[i] = x[i] >= Real(0) ? t : t - 1;
		}
		return c;
	}

	static uint64_t spread21(uint32_t v)
	{
		uint64_t x = v & 0x1fffffu;
		x = (x | x << 32) & 0x1f00000000ffffull;
		x = (x | x << 16) & 0x1f0000ff0000ffull;
		x = (x | x << 8) & 0x100f00f00f00f00full;
		x = (x | x << 4) & 0x10c30c30c30c30c3ull;
		x = (x | x << 2) & 0x1249249249249249ull;
		return x;
	}

	static uint64_t morton(CellKey const& c)
	{
		uint32_t const bias = uint32_t(INT_MIN);
		return spread21(uint32_t(c.k[0]) - bias) | spread21(uint32_t(c.k[1]) - bias) << 1 |
			spread21(uint32_t(c.k[2]) - bias) << 2;
	}

	void rebuild()
	{
		m_map.clear();
		for (unsigned s = 0; s < m_sets.size(); s++)
			for (std::size_t i = 0; i < m_sets[s].m_n; i++)
				m_map[cell_of(m_sets[s].m_x + 3 * i)].push_back(Entry{ s, unsigned(i) });
		m_ready = true;
	}

	Real m_r, m_r2, m_inv;
	std::vector<PointSet> m_sets;
	std::unordered_map<CellKey, std::vector<Entry>, CellKeyHasher> m_map;
	bool m_ready = false;
};
}  // namespace CompactNSearch
