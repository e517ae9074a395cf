"""CPU oracle for the trial windowing of the reference's project/segment.py.
TEST INFRASTRUCTURE ONLY.

    transition_indices   restates _transition_indices (segment.py:667-755) literally: the
                         alternating search over xor / and masks, 40 times, restarting each
                         search AT the index just found, windows cut at the end of the signal
    organize             restates _organize_transitions (segment.py:787-917)
    window_rows          restates DeviceData.__getitem__(slice) (user_data.py:727-731 with the
                         index maps of :626-661): rows [to_index(start), to_index(stop))

Parity pinning: the reference has no test for segment.py, so this oracle is pinned by
tests/golden/segment_D.json, produced by running the reference's own Segmenter on a synthetic
trial in the build container (oracle/make_golden.py).
"""
import numpy as np


def transition_indices(left, right, min_phase_size=10, num_segments=40):
    left = np.asarray(left, dtype=np.float64)
    right = np.asarray(right, dtype=np.float64)
    a, b = left != 0, right != 0
    masks = {1: np.logical_xor(a, b), 2: np.logical_and(a, b)}
    n = len(left)
    out = []
    cursor = 0
    legs = 1
    while len(out) < num_segments or num_segments == 0:
        mask = masks[legs][cursor:]
        found = None
        for ind in np.where(mask)[0]:
            if mask[ind : ind + min_phase_size].all():
                found = int(ind)
                break
        if found is None:
            if num_segments == 0:
                return out
            raise ValueError("fewer transitions than requested")
        cursor += found
        out.append(cursor)
        legs = 3 - legs
        if cursor >= n:
            break
    return out


def to_framesubfr(index, num_subframes):
    return index // num_subframes + 1, index % num_subframes


def organize(transitions, left, right, num_subframes):
    """Returns a list of dicts (trecho, cycle, order, phase, start, stop) in trecho/cycle/phase order."""
    out = []
    for n in range(4):
        starts = list(transitions[10 * n + 1 : 10 * n + 9])
        end = transitions[10 * n + 9]
        ind = starts[1]
        l, r = left[ind] != 0, right[ind] != 0
        if l == r:
            raise ValueError("expected exactly one loaded plate")
        second = "BL" if l else "AS"
        if n % 2 == 0:
            names = ["DAA", "BL", "DAE", "AS"] if second == "BL" else ["DAE", "AS", "DAA", "BL"]
        else:
            names = ["DAE", "BL", "DAA", "AS"] if second == "BL" else ["DAA", "AS", "DAE", "BL"]
        for c, bounds in enumerate((starts[:5], starts[4:] + [end])):
            for i in range(4):
                out.append(
                    {
                        "trecho": n, "cycle": c, "order": i + 1, "phase": names[i],
                        "start": to_framesubfr(bounds[i], num_subframes),
                        "stop": to_framesubfr(bounds[i + 1] - 1, num_subframes),
                    }
                )
    return out


def window_rows(section, start, stop, num_subframes):
    """Row range [a, b) that device[slice(start, stop)] selects (section 1 = forces/EMG, 2 = markers)."""
    if section == 1:
        return (start[0] - 1) * num_subframes + start[1], (stop[0] - 1) * num_subframes + stop[1]
    return start[0] - 1, stop[0] - 1
