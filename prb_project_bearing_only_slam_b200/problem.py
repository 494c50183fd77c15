"""Host-side problem container: the id -> stix bookkeeping of the reference's State
(framework/state.cpp:20-67) done once, vectorised, instead of per edge per iteration."""
import numpy as np


def xyt_to_xycs(poses_xyt):
    """v2t (framework/definitions.hpp:45-53) for an [NP,3] array -> wire format [NP,4] = x, y, c, s."""
    p = np.asarray(poses_xyt, dtype=np.float64)
    out = np.empty((p.shape[0], 4))
    out[:, 0] = p[:, 0]
    out[:, 1] = p[:, 1]
    out[:, 2] = np.cos(p[:, 2])
    out[:, 3] = np.sin(p[:, 2])
    return out


class Problem:
    """Edges with ids resolved to stix.

    pose stix = order of `pose_ids` (insertion order, State::add_pose); landmark stix = ASCENDING
    landmark id over the ids seen in the bearing edges, as triangulate_landmarks inserts them
    (slam/triangulation.cpp:65-74), unless `lm_ids` is given explicitly (VERTEX_XY file order).
    Unknown ids raise KeyError like std::map::at in the reference.
    """

    def __init__(self, pose_ids, b_pose_id, b_lm_id, b_z, o_src_id, o_dst_id, o_z, o_omega, fixed_pose_id=-1,
                 lm_ids=None, b_omega=None):
        self.pose_ids = np.asarray(pose_ids, np.int32)
        self.NP = len(self.pose_ids)
        b_lm_id = np.asarray(b_lm_id, np.int32)
        self.lm_ids = np.unique(b_lm_id).astype(np.int32) if lm_ids is None else np.asarray(lm_ids, np.int32)
        self.NL = len(self.lm_ids)
        self.b_pose = self._resolve(self.pose_ids, np.asarray(b_pose_id, np.int32), "pose")
        self.b_lm = self._resolve(self.lm_ids, b_lm_id, "landmark")
        self.o_src = self._resolve(self.pose_ids, np.asarray(o_src_id, np.int32), "pose")
        self.o_dst = self._resolve(self.pose_ids, np.asarray(o_dst_id, np.int32), "pose")
        self.b_z = np.asarray(b_z, np.float64)
        self.b_omega = None if b_omega is None else np.asarray(b_omega, np.float64)
        self.o_z = np.asarray(o_z, np.float64).reshape(-1, 3)
        self.o_omega = np.asarray(o_omega, np.float64).reshape(-1, 9)
        # default_pose_id (framework/state.cpp:65-67) = id at stix 0
        self.fixed_pose_id = int(fixed_pose_id) if fixed_pose_id >= 0 else int(self.pose_ids[0])
        self.fixed_stix = int(self._resolve(self.pose_ids, np.array([self.fixed_pose_id], np.int32), "pose")[0])
        self.Eb, self.Eo = len(self.b_z), len(self.o_src)

    @staticmethod
    def _resolve(ids, query, what):
        # last insertion wins on duplicate ids, like the reference's std::map overwrite
        order = np.argsort(ids, kind="stable")
        sorted_ids = ids[order]
        pos = np.searchsorted(sorted_ids, query, side="right") - 1
        if len(query) and (np.any(pos < 0) or np.any(sorted_ids[np.clip(pos, 0, None)] != query)):
            raise KeyError("unknown %s id in an edge" % what)
        return order[pos].astype(np.int32) if len(query) else np.zeros(0, np.int32)

    def upload(self, ctx):
        ctx.upload_problem(self.NP, self.NL, self.fixed_stix, self.b_pose, self.b_lm, self.b_z, self.b_omega,
                           self.o_src, self.o_dst, self.o_z, self.o_omega)
