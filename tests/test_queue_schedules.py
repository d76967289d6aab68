"""The parallel schedules the search kernel uses on its priority queue, modelled step by step on the CPU and compared
with the sequential libstdc++ algorithms they replace (bits/stl_heap.h: __push_heap :135-148, __adjust_heap :224-267).

The queue decides which node a search expands next, and among nodes that tie on (cost, num_dets) the order falls out of
the binary heap's array layout — so a schedule must reproduce the ARRAY, not just the heap property.  These models
follow tesseract_decoder_b200/csrc/search_fast.cu line by line (phase 7, "whole heap" regime, and heap_pop) with the
kernel's lane structure; keys are drawn from a small alphabet so that ties are the common case."""
import random

LANES = 32


def greater(a, b):
    """Node::operator> (tesseract.cc:119-121): larger cost, or equal cost and FEWER detectors."""
    return a[0] > b[0] or (a[0] == b[0] and a[1] < b[1])


def push_heap_sequential(heap, value):
    """push_back + std::push_heap(first, last, std::greater<>): bits/stl_heap.h:135-148."""
    heap.append(value)
    hole = len(heap) - 1
    while hole > 0:
        parent = (hole - 1) // 2
        if not greater(heap[parent], value):
            break
        heap[hole] = heap[parent]
        hole = parent
    heap[hole] = value


def pop_heap_sequential(heap):
    """std::pop_heap + pop_back with std::greater<>: __adjust_heap (bits/stl_heap.h:224-267) followed by __push_heap."""
    top = heap[0]
    value = heap.pop()
    n = len(heap)
    if n == 0:
        return top
    hole = 0
    child = 0
    while child < (n - 1) // 2:
        child = 2 * (child + 1)
        if greater(heap[child], heap[child - 1]):
            child -= 1
        heap[hole] = heap[child]
        hole = child
    if (n & 1) == 0 and child == (n - 2) // 2:
        child = 2 * (child + 1)
        heap[hole] = heap[child - 1]
        hole = child - 1
    while hole > 0:
        parent = (hole - 1) // 2
        if not greater(heap[parent], value):
            break
        heap[hole] = heap[parent]
        hole = parent
    heap[hole] = value
    return top


def push_pipeline(heap, values, reads_before_writes):
    """search_fast.cu phase 7, regime (a): the whole heap staged, the new nodes pre-placed at their leaves, push i making
    its step s at time 2 i + s, lane l running pushes l, l + 32, ...  `reads_before_writes` models a warp in lock step
    (every lane's parent read precedes every lane's write within a step); False runs the lanes of a step one after the
    other — the outcome must not depend on it, because pushes two steps apart never touch the same tree level."""
    hs, K = len(heap), len(values)
    stage = list(heap) + list(values)  # pre-placement: stage[hs + i] = values[i]
    my = list(range(LANES))
    active = [False] * LANES
    hole1 = [0] * LANES
    value = [None] * LANES
    nextv = [stage[hs + l] if l < K else None for l in range(LANES)]
    t_end = 2 * (K - 1) + (hs + K).bit_length() + 1
    for t in range(t_end):
        for l in range(LANES):
            if not active[l] and my[l] < K and t == 2 * my[l]:
                value[l] = nextv[l]
                active[l] = True
                hole1[l] = hs + my[l] + 1  # 1-based
        writes = []
        for l in range(LANES):
            if not active[l]:
                continue
            up, pv = False, None
            if hole1[l] > 1:
                pv = stage[(hole1[l] >> 1) - 1]
                up = greater(pv, value[l])
            w = (hole1[l] - 1, pv if up else value[l])
            if reads_before_writes:
                writes.append(w)
            else:
                stage[w[0]] = w[1]
            if up:
                hole1[l] >>= 1
            else:
                active[l] = False
                my[l] += LANES
                if my[l] < K:
                    nextv[l] = stage[hs + my[l]]  # its own leaf: nobody else has touched it yet
        for i, v in writes:
            stage[i] = v
    assert not any(active) and all(m >= K for m in my[:min(K, LANES)])
    return stage


def random_node(rng, alphabet):
    return (float(rng.randrange(alphabet)), rng.randrange(3))


def random_heap(rng, n, alphabet):
    heap = []
    for _ in range(n):
        push_heap_sequential(heap, random_node(rng, alphabet))
    # a few pops in between, so that the array is not just the image of pushes
    for _ in range(min(n // 3, 5)):
        pop_heap_sequential(heap)
        push_heap_sequential(heap, random_node(rng, alphabet))
    return heap


def test_pipeline_push_reproduces_sequential_push_heap_array():
    rng = random.Random(20261020)
    cases = 0
    for trial in range(1500):
        hs = rng.choice([0, 0, 1, 2, 3, 7, 8, 31, 32, 33, 100, 255, 256, 300, 511])
        K = rng.choice([1, 2, 3, 31, 32, 33, 64, 65, 96, 130, 225, 242])
        alphabet = rng.choice([1, 2, 4, 16, 1000])  # 1: every key ties
        heap = random_heap(rng, hs, alphabet)
        values = [random_node(rng, alphabet) for _ in range(K)]
        want = list(heap)
        for v in values:
            push_heap_sequential(want, v)
        for lock_step in (True, False):
            got = push_pipeline(heap, values, lock_step)
            assert got == want, (trial, hs, K, alphabet, lock_step)
        cases += 1
    assert cases == 1500


def test_pipeline_push_then_pops_give_the_reference_order():
    """End to end on the thing that matters: the sequence of popped nodes (with ties) after a staged batch of pushes."""
    rng = random.Random(7)
    for trial in range(200):
        heap = random_heap(rng, rng.randrange(0, 200), 3)
        values = [random_node(rng, 3) for _ in range(rng.randrange(96, 243))]
        want = list(heap)
        for v in values:
            push_heap_sequential(want, v)
        got = push_pipeline(heap, values, True)
        order_want = [pop_heap_sequential(want) for _ in range(len(want))]
        order_got = [pop_heap_sequential(got) for _ in range(len(got))]
        assert order_got == order_want
        assert all(not greater(a, b) for a, b in zip(order_want, order_want[1:]))  # and it is a priority order at all


def pop_heap_subtree_fetch(heap):
    """search_fast.cu heap_pop: the hole walks down four levels per memory round trip — the 30 nodes of the subtree below
    the hole are fetched together and the walk of __adjust_heap is replayed on the fetched copies; then the sift-up of the
    moved last element.  Modelled as: read a snapshot of the 4-level subtree, decide the path from the snapshot only."""
    top = heap[0]
    value = heap.pop()
    n = len(heap)
    if n == 0:
        return top
    hole = 0
    while True:
        # snapshot of the subtree of `hole`, 4 levels below it (indices < n only)
        snap = {}
        frontier = [hole]
        for _ in range(4):
            nxt = []
            for q in frontier:
                for c in (2 * q + 1, 2 * q + 2):
                    if c < n:
                        snap[c] = heap[c]
                        nxt.append(c)
            frontier = nxt
        moved = 0
        while moved < 4:
            left, right = 2 * hole + 1, 2 * hole + 2
            if right < n:  # both children: the better one moves up (ties: the left child, stl_heap.h:234)
                child = left if greater(snap[right], snap[left]) else right
            elif left < n:  # a lone left child at the very end (stl_heap.h:240-245)
                child = left
            else:
                break
            heap[hole] = snap[child]
            hole = child
            moved += 1
        if moved < 4:
            break
    while hole > 0:
        parent = (hole - 1) // 2
        if not greater(heap[parent], value):
            break
        heap[hole] = heap[parent]
        hole = parent
    heap[hole] = value
    return top


def test_subtree_fetch_pop_reproduces_sequential_pop_heap_array():
    rng = random.Random(99)
    for trial in range(600):
        n = rng.choice([1, 2, 3, 4, 5, 15, 16, 17, 30, 31, 32, 33, 63, 64, 100, 500, 1023, 1024, 1025])
        alphabet = rng.choice([1, 2, 4, 1000])
        a = random_heap(rng, n, alphabet)
        b = list(a)
        for _ in range(min(n, 40)):
            ta = pop_heap_sequential(a)
            tb = pop_heap_subtree_fetch(b)
            assert ta == tb and a == b, (trial, n, alphabet)
