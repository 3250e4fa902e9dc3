//! `environment` over the env-pool kernels of libomok_b200 (k_env_step & co., omok-ai_b200/csrc/env_kernels.cu).
//! Same public surface as the reference's environment/src/lib.rs: `Stone`, `Turn`, `GameStatus`, and `Environment`
//! with its three public fields, which the callers read directly (src/trainer.rs:224-302, gui/src/main.rs:97-108) and
//! which are therefore refreshed from the device record after every mutation.
//! NOT COMPILED in the build image of this repository (no Rust toolchain): see ../README.md.
use omok_b200_sys as sys;
use serde::{Deserialize, Serialize};
use std::fmt::Display;

#[derive(Serialize, Deserialize, Debug, Clone, Copy, PartialEq, Eq, PartialOrd, Ord, Hash)]
pub enum Stone {
    Empty,
    Black,
    White,
}

impl Display for Stone {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        f.write_str(match self {
            Stone::Empty => "-",
            Stone::Black => "X",
            Stone::White => "O",
        })
    }
}

#[derive(Serialize, Deserialize, Debug, Clone, Copy, PartialEq, Eq, PartialOrd, Ord, Hash)]
pub enum Turn {
    Black,
    White,
}

impl Turn {
    pub fn opponent(self) -> Self {
        if self == Turn::Black {
            Turn::White
        } else {
            Turn::Black
        }
    }
}

impl Display for Turn {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        f.write_str(if *self == Turn::Black { "Black" } else { "White" })
    }
}

#[derive(Serialize, Deserialize, Debug, Clone, Copy, PartialEq, Eq, PartialOrd, Ord, Hash)]
pub enum GameStatus {
    InProgress,
    Draw,
    BlackWin,
    WhiteWin,
}

impl GameStatus {
    pub fn is_terminal(self) -> bool {
        self != GameStatus::InProgress
    }
    /// status byte of the C ABI (include/omok_b200.h: 0 InProgress, 1 Draw, 2 BlackWin, 3 WhiteWin)
    pub fn from_abi(code: i8) -> Option<Self> {
        match code {
            0 => Some(GameStatus::InProgress),
            1 => Some(GameStatus::Draw),
            2 => Some(GameStatus::BlackWin),
            3 => Some(GameStatus::WhiteWin),
            _ => None, // OMK_NONE: the cell was occupied
        }
    }
}

const CELLS: usize = Environment::BOARD_SIZE * Environment::BOARD_SIZE;

pub struct Environment {
    pub turn: Turn,
    pub legal_move_count: u16,
    pub board: [Stone; CELLS],
    slot: i32, // record of the device env pool
}

impl Environment {
    pub const BOARD_SIZE: usize = 9;
    pub const SERIAL_STONE_COUNT: usize = 5;

    pub fn new() -> Self {
        let slot = sys::with(|c| {
            let slot = c.take_env();
            sys::check(unsafe { sys::omk_env_reset(c.raw(), &slot, 1) }).unwrap();
            slot
        });
        Environment { turn: Turn::Black, legal_move_count: CELLS as u16, board: [Stone::Empty; CELLS], slot }
    }

    /// Device record -> public fields.
    fn refresh(&mut self) {
        let (mut cells, mut turn, mut legal) = ([0u8; CELLS], 0u8, 0u16);
        sys::with(|c| sys::check(unsafe { sys::omk_env_get(c.raw(), &self.slot, 1, cells.as_mut_ptr(), &mut turn, &mut legal) }).unwrap());
        for (dst, &src) in self.board.iter_mut().zip(cells.iter()) {
            *dst = match src {
                1 => Stone::Black,
                2 => Stone::White,
                _ => Stone::Empty,
            };
        }
        self.turn = if turn == 0 { Turn::Black } else { Turn::White };
        self.legal_move_count = legal;
    }

    /// Public fields -> device record (the fields are `pub`: a caller may have edited them).
    fn upload(&self) {
        let mut cells = [0u8; CELLS];
        for (dst, src) in cells.iter_mut().zip(self.board.iter()) {
            *dst = *src as u8;
        }
        let turn = self.turn as u8;
        sys::with(|c| sys::check(unsafe { sys::omk_env_set(c.raw(), &self.slot, 1, cells.as_ptr(), &turn) }).unwrap());
    }

    /// The slot id, for the batched calls of `alpha-zero` (encode_nn_input over many environments in one launch).
    pub fn slot(&self) -> i32 {
        self.slot
    }

    /// reference: environment/src/lib.rs:81-102 -- the two-plane interleave `dst[index * 2 + plane]`, plane 0 = `turn`'s stones
    pub fn encode_board(&self, turn: Turn, mut dst: impl AsMut<[f32]>) {
        let dst = dst.as_mut();
        // omk_env_encode writes the 243-float network slot from the side to move's view (Player) or the other's (Opponent)
        let mode = if turn == self.turn { sys::OMK_TURN_MODE_PLAYER } else { sys::OMK_TURN_MODE_OPPONENT };
        let mut image = [0f32; 243];
        sys::with(|c| sys::check(unsafe { sys::omk_env_encode(c.raw(), &self.slot, 1, mode, image.as_mut_ptr()) }).unwrap());
        dst[..2 * CELLS].copy_from_slice(&image[..2 * CELLS]);
    }

    /// reference: environment/src/lib.rs:104-166 -- `None` (and no mutation) when the cell is occupied
    pub fn place_stone(&mut self, index: usize) -> Option<GameStatus> {
        let (action, mut status) = (index as u8, sys::OMK_NONE);
        sys::with(|c| {
            sys::check(unsafe { sys::omk_env_step(c.raw(), &self.slot, &action, 1, &mut status, std::ptr::null_mut()) }).unwrap()
        });
        let status = GameStatus::from_abi(status)?;
        self.refresh();
        Some(status)
    }
}

impl Default for Environment {
    fn default() -> Self {
        Self::new()
    }
}

impl Clone for Environment {
    fn clone(&self) -> Self {
        let mut copy = Environment::new();
        copy.turn = self.turn;
        copy.legal_move_count = self.legal_move_count;
        copy.board = self.board;
        copy.upload();
        copy
    }
}

impl Drop for Environment {
    fn drop(&mut self) {
        sys::with(|c| c.give_env(self.slot));
    }
}
