"""`import parallax` — drop-in name for users of the reference package.
Everything lives in `parallax_b200`."""
from parallax_b200 import *          # noqa: F401,F403
from parallax_b200 import (shard, log, optim, nn, Config,  # noqa: F401
                           __version__)
