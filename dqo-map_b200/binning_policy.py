"""Host policy for the rasterizer's occlusion-aware two-phase binning (include/dqo_b200.h: front_instances /
back_instances).  Pure Python, no device access: it only looks at the status words a previous forward call reported.

Both modes produce identical results; the policy only decides how much binning work is spent.

* single phase: every (Gaussian, tile) instance is emitted and sorted (what the reference does,
  rasterizer_impl.cu:303-345).  Cost ~ R.
* two phase: the nearest Gaussians (depth-rank prefix filling `front` instance slots) are binned and blended first;
  tiles whose pixels all terminated (T < T_threshold, forward.cu:823-828) are finished, and the remaining Gaussians are
  binned only into the unfinished tiles (`back` slots).  Cost ~ front + back + a second count/scan.

The blend reports how many list entries it actually staged before early termination (DQO_ST_WALKED).  When that is a small
fraction of R (heavy occlusion: most of every tile's list is never read) the two-phase mode pays off; when nearly
everything is walked it cannot, and the policy stays in / returns to single phase.
"""

ST_NUM_RENDERED, ST_TILE_NUM, ST_OVERFLOW, ST_NUM_VISIBLE, ST_R_FRONT, ST_R_BACK, ST_WALKED, ST_UNFINISHED = range(8)

MIN_INSTANCES = 1 << 21      # below this the fixed cost of the extra launches outweighs any saving
WALKED_FRACTION = 0.35       # enter two-phase when the blend walked less than this fraction of R
GIVE_UP_FRACTION = 0.75      # leave two-phase when front + back instances exceed this fraction of R
import os

# front slots per walked entry (a global depth prefix is coarser than per-tile prefixes); env override for experiments
FRONT_OVER_WALKED = float(os.environ.get("DQO_FRONT_OVER_WALKED", "2.0"))
BACK_HEADROOM = 1.5
BACK_FLOOR = 1 << 16
RETRY_AFTER = 64             # calls to wait before probing two-phase again after giving up


def _round256(n):
    return int((int(n) + 255) // 256 * 256)


class BinningPolicy:
    def __init__(self):
        self.front = 0          # 0 = single phase
        self.back = 0
        self.cooldown = 0
        self.back_is_guess = False
        self.history = []       # (mode, R, front, back_needed) of recent calls, newest last (diagnostics)

    def plan(self, capacity):
        """(front_instances, back_instances) for the next call given the instance capacity of its buffers."""
        if self.front <= 0:
            return 0, 0
        front = min(self.front, _round256(capacity // 2))
        back = min(self.back, capacity - front)
        if front <= 0 or back <= 0:
            return 0, 0
        return front, back

    def update(self, status, used_front, used_back):
        """Digest the status words of a finished call that ran with (used_front, used_back).  Returns True when that
        call has to be repeated (overflow)."""
        R = int(status[ST_NUM_RENDERED])
        if used_front <= 0:
            self.history.append(("single", R, 0, 0))
            del self.history[:-8]
            if status[ST_OVERFLOW]:
                return True
            if self.cooldown > 0:
                self.cooldown -= 1
                return False
            walked = int(status[ST_WALKED])
            if R >= MIN_INSTANCES and walked <= WALKED_FRACTION * R:
                self.front = _round256(min(max(FRONT_OVER_WALKED * walked, R / 16.0), R / 2.0))
                self.back = max(R - self.front, BACK_FLOOR) + 1024   # safe first guess: everything could be unfinished
                self.back_is_guess = True
            return False
        r_front, r_back = int(status[ST_R_FRONT]), int(status[ST_R_BACK])
        self.history.append(("two_phase", R, used_front, r_back))
        del self.history[:-8]
        if status[ST_OVERFLOW]:
            self.back = int(r_back * BACK_HEADROOM) + BACK_FLOOR
            return True
        if R < MIN_INSTANCES or r_front + r_back > GIVE_UP_FRACTION * R:
            self.front, self.back, self.cooldown = 0, 0, RETRY_AFTER
            return False
        if r_back > r_front:   # too many tiles were left unfinished: deepen the front phase
            self.front = _round256(min(self.front * 1.5, R / 2.0))
        # after the first observation: shrink slowly, grow immediately (the keyframes of a window differ)
        want = int(r_back * BACK_HEADROOM) + BACK_FLOOR
        if self.back_is_guess or self.back < want:
            self.back = want
        else:
            self.back = max(want, int(self.back * 0.9))
        self.back_is_guess = False
        return False
