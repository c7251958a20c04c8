"""Host-side selection helpers of the ADAPT loops.

Same names, arguments and results as reference
openvqe/common_files/sorted_gradient.py (``corresponding_index`` :5-20,
``index_without_0`` :23-34, ``value_without_0`` :37-55, ``abs_sort_desc``
:72-88): they decide WHICH operator ADAPT adds, so their semantics -- exact-zero
removal, float-equality matching, ties resolved to the lowest pool index --
are reproduced exactly.  They stay on the host (O(|pool|^2) Python, negligible).
"""
from collections import Counter


def index_without_0(my_list):
    """Positions whose entry is not exactly zero."""
    return [k for k, v in enumerate(my_list) if v != 0]


def value_without_0(my_list):
    """Entries that are not exactly zero, original order."""
    return [v for v in my_list if v != 0]


def abs_sort_desc(my_list):
    """Sort by magnitude, descending, IN PLACE (the reference mutates and returns
    its argument); negative entries get their sign back on the first
    ``count`` slots holding their magnitude."""
    neg_count = Counter(v for v in my_list if v < 0)
    for k, v in enumerate(my_list):
        if v < 0:
            my_list[k] = -v
    my_list.sort(reverse=True)
    seen = set()
    for v in list(my_list):
        if v > 0 and -v in neg_count and -v not in seen:
            seen.add(-v)
            first = my_list.index(v)
            for j in range(first, min(len(my_list), first + neg_count[-v])):
                my_list[j] = -my_list[j]
    return my_list


def corresponding_index(new_list, new_list_index, sorted_new):
    """Pool indices in the order of ``sorted_new``; equal values map to ascending
    indices and every index appears once."""
    order, taken = [], set()
    for target in sorted_new:
        for value, idx in zip(new_list, new_list_index):
            if value == target and idx not in taken:
                taken.add(idx)
                order.append(idx)
    return order
