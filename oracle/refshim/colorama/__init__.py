"""Test-infrastructure stub for the one dependency of /root/reference that is not
installed here (colorama is used for ANSI colour strings only: board.py:5, unit.py:5,
structure.py:5, spell.py:5, player.py:10, utils.py:2).  Every attribute is ""."""


class _Blank:
    def __getattr__(self, _name):
        return ""


Back = _Blank()
Fore = _Blank()
Style = _Blank()
