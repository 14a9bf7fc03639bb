// LRU model of the L2 for one 40^6 tuple of config 2 ("DRAM traffic"): the sub-tiles are visited in a given
// order; each touches, for every one of the 18 descriptors (9 splits x {sd_t_d1_K, sd_t_d2_K}), one G1 and one G2 block
// column (K = 40 -> 5 stages x 4 KiB = 20 KiB).  Counts the misses of an LRU cache of `cap` MiB.
//   gcc -O2 -o l2_order_model tools/l2_order_model.c && ./l2_order_model <mode> <cap MiB> [brick edge]
//   mode 0: the kernel's linear order (h3 block fastest ... p4 block slowest)
//   mode 1: bricks of a^5 over the five fast positions, p4 block still slowest (keeps p4 slabs contiguous item ranges)
//   mode 2: bricks of a^6
// Results (GB of DRAM traffic per tuple; the measured 22.0 GB per two-tuple launch = 11 GB per tuple):
//   cap 110 MiB: mode 0: 16.2   mode 1 (a=5): 9.6   mode 2 (a=5): 4.7    (compulsory: 0.74)
//   cap  80 MiB: mode 0: 16.2   mode 1 (a=5): 11.9  mode 2 (a=5): 5.2
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#define NB 10
static int prv[40000], nxt[40000], in[40000];
static int head=-1, tail=-1; static long used=0;
static void unlink_(int b){ if(prv[b]>=0) nxt[prv[b]]=nxt[b]; else head=nxt[b]; if(nxt[b]>=0) prv[nxt[b]]=prv[b]; else tail=prv[b]; }
static void push_front(int b){ prv[b]=-1; nxt[b]=head; if(head>=0) prv[head]=b; head=b; if(tail<0) tail=b; }
int main(int argc,char**argv){
  int mode=atoi(argv[1]); long cap=atol(argv[2])*1024L*1024L/ (20*1024); // capacity in block columns
  int a=argc>3?atoi(argv[3]):5;
  // splits: pa in {3,4,5}, hb in {0,1,2}; G1 = {pa, holes != hb}; G2 = {hb, particles != pa}
  long misses=0, acc=0;
  memset(in,0,sizeof in);
  long total=1; for(int i=0;i<6;i++) total*=NB;
  for(long it=0; it<total; it++){
    int b[6]; long r=it;
    if(mode==0){ for(int q=0;q<6;q++){ b[q]=r%NB; r/=NB; } }
    else if(mode==1){ // bricks of a^5 over positions 0..4 (p4 slowest, plain): inner digits first then outer
      int inner[5], outer[5]; int nbr=(NB+a-1)/a;
      for(int q=0;q<5;q++){ inner[q]=r%a; r/=a; }
      for(int q=0;q<5;q++){ outer[q]=r%nbr; r/=nbr; }
      b[5]=r; for(int q=0;q<5;q++) b[q]=outer[q]*a+inner[q];
    } else { // mode 2: bricks of a^6
      int inner[6], outer[6]; int nbr=(NB+a-1)/a;
      for(int q=0;q<6;q++){ inner[q]=r%a; r/=a; }
      for(int q=0;q<6;q++){ outer[q]=r%nbr; r/=nbr; }
      for(int q=0;q<6;q++) b[q]=outer[q]*a+inner[q];
    }
    for(int s=0;s<9;s++){
      int pa=3+s/3, hb=s%3;
      int g1[3],g2[3],n1=0,n2=0; g1[n1++]=pa; for(int h=0;h<3;h++) if(h!=hb) g1[n1++]=h;
      g2[n2++]=hb; for(int p=3;p<6;p++) if(p!=pa) g2[n2++]=p;
      int c1=(b[g1[0]]*NB+b[g1[1]])*NB+b[g1[2]], c2=(b[g2[0]]*NB+b[g2[1]])*NB+b[g2[2]];
      for(int d=0; d<2; d++){
        int ids[2]={ ((s*2+d)*2+0)*1000+c1, ((s*2+d)*2+1)*1000+c2 };
        for(int k=0;k<2;k++){ int id=ids[k]; acc++;
          if(in[id]){ unlink_(id); push_front(id); }
          else { misses++; if(used>=cap){ int v=tail; unlink_(v); in[v]=0; used--; } in[id]=1; used++; push_front(id); }
        }
      }
    }
  }
  printf("mode %d a %d cap %ld cols: accesses %ld misses %ld -> DRAM %.2f GB per tuple (compulsory %.2f GB)\n", mode,a,cap,acc,misses, misses*20.0*1024/1e9, 36000*20.0*1024/1e9);
  return 0;
}
