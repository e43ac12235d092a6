#pragma once
// tns::NeighborList -- handle to one neighbour list, source compatible with the reference's
// TreeNSearch/source/NeighborList.h:8-39 (size(), operator[], get_ptr()).
//
// The engine stores every list exactly like the reference does (TreeNSearch.h:395): a count word followed by the ids,
// [n, j0, j1, ..., j(n-1)], inside one pinned host buffer per (set_i, set_j) pair.  The handle points at the count word.
#include <cstddef>

namespace tns
{
	class TreeNSearch;

	class NeighborList
	{
	public:
		/** Number of neighbours in the list. */
		inline int size() const { return head_[0]; }
		/** Index (local to set_j) of the i-th neighbour. */
		inline int operator[](const std::size_t i) const { return head_[1 + i]; }
		/** Pointer to the first neighbour id. */
		inline const int* get_ptr() const { return head_ + 1; }

		inline const int* begin() const { return head_ + 1; }
		inline const int* end() const { return head_ + 1 + head_[0]; }

	private:
		explicit NeighborList(const int* count_word) : head_(count_word) {}
		const int* head_;
		friend class TreeNSearch;
	};
}
