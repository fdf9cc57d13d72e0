"""`GCBC` — drop-in for `hulc.models.gcbc.GCBC` (reference hulc/models/gcbc.py:11-181): goal-conditioned behaviour cloning
ablation.  The decoder receives an empty plan (`plan_features = 0`, gcbc.py:16-48), there is no prior and no KL; the
posterior transformer still runs because its `seq_feat` feeds the CLIP auxiliary loss (gcbc.py:50-181)."""
from .hulc import Hulc


class GCBC(Hulc):
    MODEL = "gcbc"

    # ---- inference (gcbc.py:281-317): the goal is encoded once per rollout, there is no plan and no re-planning -------------------
    def reset(self):
        self.latent_goal = None
        self.engine.infer_reset()

    def step(self, obs, goal, *, sample_u=None):
        """One control step.  Unlike the reference, whose decoder keeps its hidden state across `reset()` calls (nothing clears it), a new
        rollout starts from a zero hidden state here."""
        import torch

        with torch.no_grad():
            dev = self.engine.device
            st, gr = obs["rgb_obs"]["rgb_static"].to(dev).float(), obs["rgb_obs"]["rgb_gripper"].to(dev).float()
            if getattr(self, "latent_goal", None) is None:
                if isinstance(goal, str):
                    lang = torch.from_numpy(self.lang_embeddings[goal]).to(dev).squeeze(0).float()
                    _, self.latent_goal = self.engine.infer_plan(st[0].contiguous(), gr[0].contiguous(), lang=lang.reshape(1, -1).contiguous())
                else:
                    gs, gg = goal["rgb_obs"]["rgb_static"].to(dev).float(), goal["rgb_obs"]["rgb_gripper"].to(dev).float()
                    _, self.latent_goal = self.engine.infer_plan(torch.cat([st, gs], 1)[0].contiguous(), torch.cat([gr, gg], 1)[0].contiguous())
            return self.engine.infer_act(st[0].contiguous(), gr[0].contiguous(), obs["robot_obs_raw"].to(dev).float().reshape(1, -1), sample_u=sample_u)
