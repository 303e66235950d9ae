"""
ContactManager: net contact force / mean contact position per tracked link, and feet air-time.

API of genesis_forge/managers/contact/contact_manager.py.  The reference pulls the padded contact
tensors from the collider, launches a Taichi scatter (contact/kernel.py:5-90, float atomics) and
runs the air-time state machine with eager `torch.where`s (:434-477); here all of that is part of
the fused post-physics kernel, with ordered (deterministic) accumulation.  This object resolves
link name patterns (:342-382), owns the output/state tensors and offers the helper predicates.
"""
from __future__ import annotations

import re

import torch

from .._gs import gs
from .base import BaseManager


class ContactManager(BaseManager):
    def __init__(
        self,
        env,
        link_names: list[str],
        entity_attr: str = "robot",
        with_entity_attr: str = None,
        with_links_names: list[str] = None,
        track_air_time: bool = False,
        air_time_contact_threshold: float = 1.0,
        debug_visualizer: bool = False,
        debug_visualizer_cfg: dict | None = None,
    ):
        super().__init__(env, "contact")
        self._link_names = link_names
        self._air_time_contact_threshold = air_time_contact_threshold
        self._track_air_time = track_air_time
        self._entity_attr = entity_attr
        self._link_ids = None
        self._local_link_ids = None
        self._with_entity_attr = with_entity_attr
        self._with_links_names = with_links_names
        self._with_link_ids = torch.empty(0, device=gs.device)
        self._with_local_link_ids = None
        self._has_with_filter = with_entity_attr is not None or with_links_names is not None
        self.debug_visualizer = debug_visualizer
        self.visualizer_cfg = dict(debug_visualizer_cfg or {})
        self.contacts: torch.Tensor | None = None
        self.contact_positions: torch.Tensor | None = None
        self._air: torch.Tensor | None = None  # (4, N, Lc): last_air, cur_air, last_contact, cur_contact

    # -- state views ----------------------------------------------------------------------------
    @property
    def last_air_time(self):
        return None if self._air is None else self._air[0]

    @property
    def current_air_time(self):
        return None if self._air is None else self._air[1]

    @property
    def last_contact_time(self):
        return None if self._air is None else self._air[2]

    @property
    def current_contact_time(self):
        return None if self._air is None else self._air[3]

    @property
    def link_ids(self) -> torch.Tensor:
        return self._link_ids

    @property
    def local_link_ids(self) -> torch.Tensor:
        return self._local_link_ids

    # -- helpers --------------------------------------------------------------------------------
    def _require_air_time(self):
        if not self._track_air_time:
            raise RuntimeError(
                "The contact manager is not configured to track air time. "
                "Please enable 'track_air_time' in the manager configuration."
            )

    def has_made_contact(self, dt: float, time_margin: float = 1.0e-8) -> torch.Tensor:
        """Links that established contact within the last `dt` seconds (contact_manager.py:198-224)."""
        self._require_air_time()
        t = self.current_contact_time
        return (t > 0.0) * (t < (dt + time_margin))

    def has_broken_contact(self, dt: float, time_margin: float = 1.0e-8) -> torch.Tensor:
        """Links that broke contact within the last `dt` seconds (contact_manager.py:226-256)."""
        self._require_air_time()
        t = self.current_air_time
        return (t > 0.0) * (t < (dt + time_margin))

    def get_contact_forces(self, link_idx: int) -> torch.Tensor:
        idx = torch.nonzero(self._link_ids == link_idx)[0]
        return self.contacts[:, idx, :]

    # -- operations -----------------------------------------------------------------------------
    def build(self):
        super().build()
        self._link_ids, self._local_link_ids = self._get_links_idx(self._entity_attr, self._link_names)
        if self._with_entity_attr or self._with_links_names:
            with_attr = self._with_entity_attr if self._with_entity_attr is not None else "robot"
            self._with_link_ids, self._with_local_link_ids = self._get_links_idx(with_attr, self._with_links_names)
        n, lc = self.env.num_envs, self._link_ids.shape[0]
        self.contacts = torch.zeros((n, lc, 3), device=gs.device)
        self.contact_positions = torch.zeros((n, lc, 3), device=gs.device)
        if self._track_air_time:
            self._air = torch.zeros((4, n, lc), device=gs.device)

    def reset(self, envs_idx=None):
        """Host-side variant of contact_manager.py:316-329 (the per-step reset is in-kernel)."""
        if not self.enabled or not self._track_air_time:
            return
        if envs_idx is None:
            self._air.zero_()
        else:
            self._air[:, envs_idx] = 0.0

    def _get_links_idx(self, entity_attr: str, names: list[str] = None):
        entity = getattr(self.env, entity_attr)
        ids, local_ids = [], []
        if names is None:
            for link in entity.links:
                ids.append(link.idx)
                local_ids.append(link.idx_local)
        else:
            for pattern in names:
                found = False
                for link in entity.links:
                    if pattern == link.name or re.match(f"^{pattern}$", link.name):
                        ids.append(link.idx)
                        local_ids.append(link.idx_local)
                        found = True
                if not found:
                    available = [link.name for link in entity.links]
                    raise RuntimeError(
                        f"Link '{pattern}' not found in entity '{self._entity_attr}'.\nAvailable links: {available}"
                    )
        return torch.tensor(ids, device=gs.device), torch.tensor(local_ids, device=gs.device)

    def __repr__(self):
        attrs = [f"link_names={self._link_names}", f"entity_attr={self._entity_attr}"]
        if self._with_entity_attr:
            attrs.append(f"with_entity_attr={self._with_entity_attr}")
        if self._with_links_names:
            attrs.append(f"with_links_names={self._with_links_names}")
        if self._track_air_time:
            attrs.append(f"track_air_time=True, air_time_contact_threshold={self._air_time_contact_threshold}")
        return f"{self.__class__.__name__}({', '.join(attrs)})"
