"""EngineContext: owns an Engine and its keys.  Mirror of
/root/reference/engine_context.py:9-85 (same constructor signature switch, same attribute
names), bound to this package's Engine instead of ``desilofhe.Engine``."""
from __future__ import annotations

from typing import Optional, Sequence

from ..engine import Engine


class EngineContext:
    def __init__(self, signature: int, *, max_level: int = 30, mode: str = 'cpu',
                 use_bootstrap: bool = True, use_multiparty: bool = False, thread_count: int = 0,
                 device_id: int = 0, fixed_rotation: bool = False, delta_list: Optional[Sequence[int]] = None,
                 log_coeff_count: int = 0, special_prime_count: int = 0,
                 _engine_kwargs: Optional[dict] = None, rotation_steps: Optional[Sequence[int]] = None) -> None:
        extra = dict(_engine_kwargs or {})
        if signature == 1:
            self.engine = Engine(mode=mode, use_bootstrap=use_bootstrap, use_multiparty=use_multiparty,
                                 thread_count=thread_count, device_id=device_id, **extra)
        elif signature == 2:
            self.engine = Engine(max_level=max_level, mode=mode, use_multiparty=use_multiparty,
                                 thread_count=thread_count, device_id=device_id, **extra)
        elif signature == 3:
            self.engine = Engine(log_coeff_count=log_coeff_count, special_prime_count=special_prime_count,
                                 mode=mode, use_multiparty=use_multiparty, thread_count=thread_count,
                                 device_id=device_id, **extra)
        else:
            raise ValueError(f"Unsupported signature: {signature}")

        self.fixed_rotation_key_list = []
        eng = self.engine
        self.secret_key = eng.create_secret_key()
        self.public_key = eng.create_public_key(self.secret_key)
        self.relinearization_key = eng.create_relinearization_key(self.secret_key)
        self.conjugation_key = eng.create_conjugation_key(self.secret_key)
        if rotation_steps is None:
            self.rotation_key = eng.create_rotation_key(self.secret_key)
        else:
            self.rotation_key = eng.create_rotation_key(self.secret_key, steps=list(rotation_steps))
        if fixed_rotation and delta_list is not None:
            for delta in delta_list:
                self.fixed_rotation_key_list.append(eng.create_fixed_rotation_key(self.secret_key, delta))
        self.small_bootstrap_key = eng.create_small_bootstrap_key(self.secret_key)
        self.bootstrap_key = eng.create_bootstrap_key(self.secret_key)

    def __repr__(self) -> str:  # pragma: no cover
        return f"FHEContext(engine=Engine(slot_count={self.engine.slot_count}), keys=[sk, pk, rlk, cjk, rot])"

    def encrypt(self, data):
        return self.engine.encrypt(data, self.public_key)

    def decrypt(self, ct):
        return self.engine.decrypt(ct, self.secret_key)
