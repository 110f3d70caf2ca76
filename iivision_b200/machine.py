"""Player virtual machine state as seen by the byte emitter; API-compatible with the
reference's ``machine.Machine``.  The only behaviour is ``emit``: an opcode's address
bytes (opcodes without an address, like the header, contribute none) followed by its
payload, after which the opcode may update the machine (a no-op hook upstream)."""

import itertools
from typing import Iterator


class Machine:
    def emit(self, opcode) -> Iterator[int]:
        address = opcode.emit_command(opcode) or ()
        payload = opcode.emit_data() or ()
        yield from itertools.chain(address, payload)
        opcode.apply(self)
